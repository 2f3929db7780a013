"""GPU, >= 2 devices: the row-partitioned operator at a production-like shape, one process per GPU (torchrun,
NCCL + NVLink peer memory) — tools/dist_check.py with its `--production` case: 32 planes of 256^2 rows per rank.
Product: every bit against the closed form on every rank, through b200sp_spmv_dist and through host buffers
(b200sp_spmv_dist_host); CG: same iteration count as the oracle CG of the whole operator, residual history within
1e-10 (only the grouping of the dot products differs); skew stress; graph all-gather; no spin time-outs.
Skipped on a single-GPU box (there bench.py's `parity` object carries the same checks at every N the driver runs)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_partitioned_operator_production_shape():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py"), "--production"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert p.returncode == 0 and lines, (p.returncode, p.stdout[-2000:], p.stderr[-3000:])
    out = json.loads(lines[-1])
    assert out["ok"] and out["world"] == n
    prod = out["production"]
    assert prod["planes_per_rank"] == 32
    assert all(v == 1 for k, v in prod.items() if k.endswith("_exact")), prod
    assert prod["cg_hist_ok"] == 1 and prod["cg_hist_max_rel_dev"] <= 1e-10 and prod["cg_x_ok"] == 1
    assert out["skew_stress_ok"] == 1 and out["graph_gather_ok"] == 1 and out["comm_timeouts"] == 0
