"""Timeline of b200sp_spmv_host's chunk pipeline on the headline operator (B200SP_HOST_TRACE=1 -> stderr)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200SP_HOST_TRACE"] = "1"
import torch

import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import gallery

h = cusp.default_handle()
A = gallery.poisson("dia", 7, (256, 256, 256), dtype=torch.float64)
x = torch.rand(A.num_cols, dtype=torch.float64).pin_memory()
y = torch.empty(A.num_rows, dtype=torch.float64).pin_memory()
d = A.descriptor()
for i in range(3):
    t0 = time.perf_counter()
    h.spmv_host(d, x, y)
    print(f"call {i}: {(time.perf_counter() - t0) * 1e3:.3f} ms wall", file=sys.stderr)
