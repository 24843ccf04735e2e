"""One process per GPU: sharding of the MSM path and the single collective that recombines it.

  * single large MSM   -> contiguous POINT RANGES per rank; each rank reduces its slice to one affine partial;
                          one all-gather of 96 B per rank; every rank sums the partials (tb200_g1_sum).
  * row commitments    -> ROWS per rank over the replicated SRS; one all-gather of rows/G x 96 B; no reduction.
  * pairing product t  -> PAIRS per rank (the row commitments a rank computed, with its slice of h_vec): each rank reduces
                          its pairs to one partial Miller product (576 B, no final exponentiation); one all-gather;
                          every rank multiplies the partials and applies the single final exponentiation.
  * MIPP rounds        -> replicas only (log-depth serial chain over <= 2^13 points; not worth sharding).

The data path has no other communication. `local_msm`, `local_rows` and `sum_points` default to the CUDA engine; the
CPU test-suite (gloo, world_size 2) injects stand-ins to exercise exactly this host logic without a GPU.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of `n` units for `rank`: sizes differ by at most one, earlier ranks take the surplus."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _engine_sum(points: np.ndarray) -> np.ndarray:
    from . import msm

    return msm.g1_sum(points)


def _engine_msm(bases: np.ndarray, scalars: np.ndarray, mont: bool) -> np.ndarray:
    from . import msm

    return msm.msm_unchecked(bases, scalars) if mont else msm.msm_bigint(bases, scalars)


def _device() -> torch.device:
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_gather_points(local: np.ndarray, group=None) -> np.ndarray:
    """All-gather of equally sized [k, 12] uint64 point blocks (the only collective on the path)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local = np.ascontiguousarray(local, dtype=np.uint64).reshape(-1, 12)
    if world == 1:
        return local.copy()
    dev = _device()
    mine = torch.from_numpy(local.view(np.int64)).to(dev)
    out = torch.empty((world * local.shape[0], 12), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.cpu().numpy().view(np.uint64)


def msm_sharded(bases_local: np.ndarray, scalars_local: np.ndarray, mont: bool = False, group=None,
                local_msm: Optional[Callable] = None, sum_points: Optional[Callable] = None) -> np.ndarray:
    """MSM over the union of every rank's slice. Each rank passes ITS slice (see shard_range); all ranks return the
    same affine point."""
    local_msm = local_msm or _engine_msm
    sum_points = sum_points or _engine_sum
    partial = np.asarray(local_msm(bases_local, scalars_local, mont), dtype=np.uint64).reshape(1, 12)
    parts = all_gather_points(partial, group)
    return np.asarray(sum_points(parts), dtype=np.uint64).reshape(12)


def commit_rows_sharded(local_rows_fn: Callable[[int, int], np.ndarray], total_rows: int, group=None) -> np.ndarray:
    """Row commitments sharded by row: `local_rows_fn(lo, hi)` returns the [hi-lo, 12] commitments of this rank's rows
    (e.g. a tb200_msm_g1_batch call on the column range Z[:, lo:hi] of the sqrt matrix). Returns all rows in order on
    every rank. Ranks are padded to the largest shard so a single fixed-size all-gather suffices."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(total_rows, rank, world)
    mine = np.asarray(local_rows_fn(lo, hi), dtype=np.uint64).reshape(hi - lo, 12)
    width = -(-total_rows // world)
    padded = np.zeros((width, 12), dtype=np.uint64)
    padded[: hi - lo] = mine
    allp = all_gather_points(padded, group).reshape(world, width, 12)
    out = np.zeros((total_rows, 12), dtype=np.uint64)
    for r in range(world):
        rlo, rhi = shard_range(total_rows, r, world)
        out[rlo:rhi] = allp[r, : rhi - rlo]
    return out


def all_gather_words(local: np.ndarray, group=None) -> np.ndarray:
    """All-gather of one equally sized uint64 vector per rank -> [world, len]."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local = np.ascontiguousarray(local, dtype=np.uint64).reshape(-1)
    if world == 1:
        return local.reshape(1, -1).copy()
    dev = _device()
    mine = torch.from_numpy(local.view(np.int64)).to(dev)
    out = torch.empty((world * local.shape[0],), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.cpu().numpy().view(np.uint64).reshape(world, -1)


def _engine_miller_product(g1s: np.ndarray, g2s: np.ndarray) -> np.ndarray:
    from . import pairing

    return pairing.miller_product(g1s, g2s)


def _engine_combine(parts: np.ndarray) -> np.ndarray:
    from . import pairing

    return pairing.final_exponentiation_of_product(parts)


def multi_pairing_sharded(g1_local: np.ndarray, g2_local: np.ndarray, group=None,
                          local_miller_product: Optional[Callable] = None,
                          combine: Optional[Callable] = None) -> np.ndarray:
    """`E::multi_pairing` over the union of every rank's pairs (e.g. `t = multi_pairing(comm_list, h_vec)` of
    `Polynomial::commit`, src/sqrt_pst.rs:131-144, with the rows sharded as in commit_rows_sharded): one partial Miller
    product per rank, one all-gather of 576 B per rank, one final exponentiation on every rank. All ranks return the same
    GT element."""
    local_miller_product = local_miller_product or _engine_miller_product
    combine = combine or _engine_combine
    partial = np.asarray(local_miller_product(g1_local, g2_local), dtype=np.uint64).reshape(72)
    parts = all_gather_words(partial, group)
    return np.asarray(combine(parts), dtype=np.uint64).reshape(72)
