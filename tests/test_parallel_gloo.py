"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: sharding + the single all-gather + recombination.
The local MSM / point sum are stood in by the oracle so that exactly testudo_b200/parallel.py is exercised."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as h
from oracle import bls12_377 as o
from testudo_b200 import parallel


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1000, 1 << 13):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_msm(bases, scalars, mont):
    from oracle import cpu

    return cpu.msm_g1(bases, scalars, mont=mont)


def _oracle_sum(points):
    acc = None
    for row in points:
        acc = o.add(acc, h.pt_from_np(row))
    return h.pts_to_np([acc])[0]


def _worker(rank, world, port, n, rows, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pts, dl = o.rand_points(n, 11)
        sc = o.rand_scalars(n, 12)
        B, S = h.pts_to_np(pts), h.scalars_to_np(sc)
        lo, hi = parallel.shard_range(n, rank, world)
        got = parallel.msm_sharded(B[lo:hi], S[lo:hi], mont=False, local_msm=_oracle_msm, sum_points=_oracle_sum)
        ok_msm = h.pt_from_np(got) == o.msm_by_dlog(dl, sc)

        srs, _ = o.rand_points(8, 13)
        z = o.rand_scalars(rows * 8, 14)

        def local_rows(rlo, rhi):
            out = []
            for i in range(rlo, rhi):
                out.append(o.msm_naive(srs, [z[j * rows + i] for j in range(8)]))
            return h.pts_to_np(out) if out else np.zeros((0, 12), dtype=np.uint64)

        allrows = parallel.commit_rows_sharded(local_rows, rows)
        exp = h.pts_to_np([o.msm_naive(srs, [z[j * rows + i] for j in range(8)]) for i in range(rows)])
        q.put((rank, ok_msm, bool(np.array_equal(allrows, exp))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,rows", [(37, 5), (64, 8)])
def test_sharded_msm_and_rows_world2(n, rows):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] and r[2] for r in results), results


def _oracle_miller_product(g1s, g2s):
    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr

    ps = [h.pt_from_np(r) for r in np.asarray(g1s, dtype=np.uint64).reshape(-1, 12)]
    qs = [o2.affine_from_words(r) for r in np.asarray(g2s, dtype=np.uint64).reshape(-1, 24)]
    return np.array(pr.to_words(pr.multi_miller_loop(ps, qs)), dtype=np.uint64)


def _oracle_combine(parts):
    from oracle import pairing as pr

    f = pr.F12_ONE
    for row in np.asarray(parts, dtype=np.uint64).reshape(-1, 72):
        f = pr.f12_mul(f, pr.from_words(row))
    return np.array(pr.to_words(pr.final_exponentiation(f)), dtype=np.uint64)


def _pairing_worker(rank, world, port, n, q):
    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ps, _ = o.rand_points(n, 21)
        qs, _ = o2.rand_points(n, 22)
        A = h.pts_to_np(ps)
        B = np.array([o2.affine_to_words(p) for p in qs], dtype=np.uint64).reshape(-1, 24)
        lo, hi = parallel.shard_range(n, rank, world)
        got = parallel.multi_pairing_sharded(A[lo:hi], B[lo:hi], local_miller_product=_oracle_miller_product,
                                             combine=_oracle_combine)
        q.put((rank, pr.from_words(got) == pr.multi_pairing(ps, qs)))
    finally:
        dist.destroy_process_group()


def test_sharded_pairing_product_world2():
    """t = multi_pairing(comm_list, h_vec) with the pairs sharded across ranks: one partial Miller product per rank, one
    all-gather of 576 B, one final exponentiation -- the host logic of parallel.multi_pairing_sharded under gloo."""
    world, n = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pairing_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1] and all(r[1] for r in results), results
