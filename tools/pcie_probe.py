import torch, time
n = 134217728
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
def chunks(k):
    c = n // k
    def f():
        for i in range(k):
            with torch.cuda.stream(s1): d1[i*c:(i+1)*c].copy_(h1[i*c:(i+1)*c], non_blocking=True)
            with torch.cuda.stream(s2): h2[i*c:(i+1)*c].copy_(d2[i*c:(i+1)*c], non_blocking=True)
    return f
for name, fn in (("h2d 134MB", h2d), ("d2h 134MB", d2h), ("both concurrently", both), ("both, 16 chunks each", chunks(16))):
    ms = t(fn); print(f"{name}: {ms:.3f} ms = {n/ms/1e6:.1f} GB/s per direction")


def pipeline(k, with_kernel):
    """the dependency pattern of b200sp_spmv_host: all H2D pieces queued at once, D2H chunk c waits for H2D piece c+1
    (through an event on a third stream that optionally runs a small kernel in between)"""
    c = n // k
    s3 = torch.cuda.Stream()
    scratch = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")

    def f():
        evs = []
        for i in range(k):
            with torch.cuda.stream(s1):
                d1[i * c:(i + 1) * c].copy_(h1[i * c:(i + 1) * c], non_blocking=True)
                e = torch.cuda.Event()
                e.record(s1)
                evs.append(e)
        for i in range(k):
            s3.wait_event(evs[min(i + 1, k - 1)])
            with torch.cuda.stream(s3):
                if with_kernel:
                    scratch.add_(1)  # 128 MB of HBM traffic, ~25 us: stands in for the chunk product
                e = torch.cuda.Event()
                e.record(s3)
            s2.wait_event(e)
            with torch.cuda.stream(s2):
                h2[i * c:(i + 1) * c].copy_(d2[i * c:(i + 1) * c], non_blocking=True)
    return f


for name, fn in (("pipeline pattern, 16 chunks, no kernel", pipeline(16, False)),
                 ("pipeline pattern, 16 chunks, kernel between", pipeline(16, True)),
                 ("pipeline pattern, 32 chunks, kernel between", pipeline(32, True))):
    ms = t(fn)
    print(f"{name}: {ms:.3f} ms = {n/ms/1e6:.1f} GB/s per direction")


def dma_fed(k, ctas, lag=1):
    """one resident kernel fed by the copy engine (tools/probe_kernels.cu: dma_fed_kernel): H2D pieces + 4-byte flag
    copies on one stream, the kernel spins on the flags and writes y straight to mapped host memory"""
    import ctypes, os
    lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "build", "libprobe.so"))
    lib.probe_dma_fed.restype = ctypes.c_int
    lib.probe_dma_fed.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    nd = n // 8
    xh = torch.rand(nd, dtype=torch.float64).pin_memory()
    yh = torch.zeros(nd, dtype=torch.float64).pin_memory()
    xd = torch.empty(nd, dtype=torch.float64, device="cuda")
    flags = torch.zeros(k, dtype=torch.int32, device="cuda")
    epochs = torch.arange(1, 1025, dtype=torch.int32).pin_memory()  # source words for the flag copies
    c = (nd + k - 1) // k
    state = {"epoch": 0}

    def f():
        state["epoch"] += 1
        ep = state["epoch"]
        rc = lib.probe_dma_fed(xd.data_ptr(), yh.data_ptr(), nd, k, lag, flags.data_ptr(), ep, ctas, s2.cuda_stream)
        assert rc == 0
        with torch.cuda.stream(s1):
            for i in range(k):
                xd[i * c:(i + 1) * c].copy_(xh[i * c:(i + 1) * c], non_blocking=True)
                flags[i:i + 1].copy_(epochs[ep - 1:ep], non_blocking=True)
    f()
    torch.cuda.synchronize()
    assert torch.equal(yh, 2 * xh), "dma-fed kernel produced a wrong y"
    return f


for k, ctas in ((16, 296), (16, 592), (32, 592), (64, 592), (16, 148)):
    ms = t(dma_fed(k, ctas), reps=8)
    print(f"dma-fed resident kernel, {k} chunks, {ctas} CTAs: {ms:.3f} ms = {n/ms/1e6:.1f} GB/s per direction")
