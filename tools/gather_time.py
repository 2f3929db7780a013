"""Times the all-gather of b200sp_spmv_dist_gather alone (empty operator, x of 2^24 fp32 in equal slices):
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/gather_time.py
B200SP_GATHER_PULL=1 selects the pull protocol."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as td
import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import dist
from cusp_autotuned_b200.matrix import coo_matrix
from cusp_autotuned_b200.partition import row_block_offsets
rank, world, local = dist.init_process_group_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
h = cusp.default_handle()
dist.init_engine_comm(h, rank, world)
n = 1 << 24
offs = row_block_offsets(n, world)
e = torch.zeros(0, dtype=torch.int32, device=dev)
A = coo_matrix(8, n, e, e, torch.zeros(0, dtype=torch.float32, device=dev))
xg = torch.arange(n, device=dev, dtype=torch.float32)
xf = torch.zeros(n, dtype=torch.float32, device=dev)
xf[offs[rank]:offs[rank+1]] = xg[offs[rank]:offs[rank+1]]
y = torch.empty(8, dtype=torch.float32, device=dev)
d = A.descriptor()
for _ in range(5): h.spmv_dist_gather(d, offs, xf, y)
ok = torch.equal(xf, xg)
torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): h.spmv_dist_gather(d, offs, xf, y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
t = torch.tensor([ms], device=dev); td.all_reduce(t, op=td.ReduceOp.MAX)
if rank == 0:
    recv = (n - (offs[1] - offs[0])) * 4
    print(json.dumps({"world": world, "protocol": "pull" if os.environ.get("B200SP_GATHER_PULL") else "push", "ok": ok,
                      "ms": float(t.item()), "recv_GBs": recv / float(t.item()) / 1e6, "timeouts": h.comm_timeouts()}), flush=True)
h.comm_destroy(); td.destroy_process_group()
