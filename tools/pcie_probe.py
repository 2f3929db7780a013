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
