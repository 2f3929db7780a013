"""The ring-stage release hazard (common.cuh: consume_before_release) reproduced in isolation, with the candidate cures:
   mode 0 arrive right after the shared-memory loads | 1 fence.proxy.async.shared::cta in between | 2 the library's
   compare-and-branch on the loaded registers | 3 fence.acq_rel.cta (membar.cta) in between.
Counts values that came back from a stage the producer had already refilled.  python tools/release_order_probe.py"""
import ctypes
import json
import os
import sys

import torch

lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "build", "libprobe.so"))
lib.probe_release_order.restype = ctypes.c_int
lib.probe_release_order.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
dev = torch.device("cuda", 0)
TILE, period = 2048, 4096
src = torch.arange(period, dtype=torch.int32, device=dev).repeat_interleave(TILE).contiguous()
tiles = 148 * 4 * 300
launches = int(sys.argv[1]) if len(sys.argv) > 1 else 40
out = {}
names = {0: "arrive right after the loads", 1: "fence.proxy.async.shared::cta before the arrive",
         2: "compare-and-branch on the loaded registers (library)", 3: "fence.acq_rel.cta before the arrive"}
for stages in (2, 3):
    for ctas_per_sm in (2, 4):
        for mode in (0, 1, 2, 3):
            bad = torch.zeros(2, dtype=torch.int64, device=dev)
            hit = 0
            for _ in range(launches):
                before = int(bad[0].item())
                rc = lib.probe_release_order(mode, tiles, period, stages, 148 * ctas_per_sm, src.data_ptr(), bad.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)
                assert rc == 0, rc
                torch.cuda.synchronize()
                hit += int(bad[0].item()) > before
            key = f"stages{stages}_ctas{ctas_per_sm}_mode{mode}"
            out[key] = {"what": names[mode], "stale_values": int(bad[0].item()), "launches_with_stale": hit, "launches": launches,
                        "values_checked_per_launch": tiles * TILE}
            print(f"## {key}: {names[mode]}: {int(bad[0].item())} stale values in {hit} of {launches} launches", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/release_order_probe.json", "w"), indent=1)
