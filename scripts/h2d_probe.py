"""Concurrent host->device bandwidth per GPU (diagnostic for the end-to-end scaling of the host-facing calls):
who limits 8 GPUs pulling pinned host memory at once -- PCIe, the host memory system, or the inter-socket link?
Usage: python scripts/h2d_probe.py [MiB per GPU]"""
import ctypes, sys, time
import torch  # loads libcudart

rt = ctypes.CDLL("libcudart.so.12")
MiB = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = MiB << 20
ndev = torch.cuda.device_count()
def ck(rc, what=""):
    if rc: raise RuntimeError(f"cuda error {rc} {what}")
def host_alloc(size, flags):
    p = ctypes.c_void_p(); ck(rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(size), ctypes.c_uint(flags)), "hostalloc"); return p
dev, streams, ev0, ev1, streams2 = [], [], [], [], []
for d in range(ndev):
    ck(rt.cudaSetDevice(d))
    p = ctypes.c_void_p(); ck(rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(n))); dev.append(p)
    s = ctypes.c_void_p(); ck(rt.cudaStreamCreateWithFlags(ctypes.byref(s), 1)); streams.append(s)
    s = ctypes.c_void_p(); ck(rt.cudaStreamCreateWithFlags(ctypes.byref(s), 1)); streams2.append(s)
    a = ctypes.c_void_p(); b = ctypes.c_void_p(); ck(rt.cudaEventCreate(ctypes.byref(a))); ck(rt.cudaEventCreate(ctypes.byref(b)))
    ev0.append(a); ev1.append(b)
def run(name, bufs, devs, split=False):
    for rep in range(2):
        t0 = time.perf_counter()
        for d in devs:
            ck(rt.cudaSetDevice(d)); ck(rt.cudaEventRecord(ev0[d], streams[d]))
            if split:
                h = n // 2
                ck(rt.cudaMemcpyAsync(dev[d], bufs[d], ctypes.c_size_t(h), 1, streams[d]))
                ck(rt.cudaMemcpyAsync(ctypes.c_void_p(dev[d].value + h), ctypes.c_void_p(bufs[d].value + h), ctypes.c_size_t(h), 1, streams2[d]))
            else:
                ck(rt.cudaMemcpyAsync(dev[d], bufs[d], ctypes.c_size_t(n), 1, streams[d]))
            ck(rt.cudaEventRecord(ev1[d], streams[d]))
        for d in devs:
            ck(rt.cudaSetDevice(d)); ck(rt.cudaStreamSynchronize(streams[d])); ck(rt.cudaStreamSynchronize(streams2[d]))
        wall = time.perf_counter() - t0
    per = []
    for d in devs:
        ms = ctypes.c_float(); ck(rt.cudaEventElapsedTime(ctypes.byref(ms), ev0[d], ev1[d])); per.append(n / ms.value / 1e6)
    print(f"{name:44s} wall {wall*1e3:7.1f} ms  aggregate {len(devs)*n/wall/1e9:6.1f} GB/s  per-GPU GB/s: " + " ".join(f"{x:5.1f}" for x in per), flush=True)
plain = [host_alloc(n, 1) for _ in range(ndev)]          # portable
for b in plain: ctypes.memset(b, 1, n)
run("plain pinned, all GPUs at once", plain, list(range(ndev)))
for d in range(ndev): run(f"plain pinned, GPU {d} alone", plain, [d])
if ndev >= 8:
    run("plain pinned, GPUs 0-3", plain, [0, 1, 2, 3]); run("plain pinned, GPUs 4-7", plain, [4, 5, 6, 7])
    run("plain pinned, GPUs 0,2,4,6", plain, [0, 2, 4, 6])
    # every GPU reads the buffer allocated for ANOTHER GPU (does placement matter?)
    run("plain pinned, rotated buffers", plain[4:] + plain[:4], list(range(ndev)))
run("plain pinned, 2 streams per GPU", plain, list(range(ndev)), split=True)
for b in plain: ck(rt.cudaFreeHost(b))
wc = [host_alloc(n, 1 | 4) for _ in range(ndev)]         # portable | write-combined
for b in wc: ctypes.memset(b, 1, n)
run("write-combined pinned, all GPUs at once", wc, list(range(ndev)))
run("write-combined pinned, GPU 0 alone", wc, [0])
