"""SASS opcode histogram per kernel of libb200sp.so (cuobjdump -sass): the instructions that show what a kernel is
built from — UBLKCP (cp.async.bulk, TMA engine), SYNCS.* (mbarrier), LDG.* by width / cache policy, LDS / STS, SHFL,
BAR, ACQBULK / griddepcontrol.  usage: python tools/sass_hist.py [regex-of-kernel-names] > profiles/rNN_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cusp_autotuned_b200", "libb200sp.so")
PAT = re.compile(sys.argv[1]) if len(sys.argv) > 1 else re.compile(
    r"dia_bulk_kernel<double, 128, 2, (false|true)>|cg_small_csr_kernel<double>|cg_update_push_p2p_kernel<double>|"
    r"ell_bulk_kernel<double, 256, 1>|csr_ring_kernel<double, 256, 8>|coo_ring_kernel<double, 512, 7>|"
    r"coo_warp_kernel<float, 256, 4, 8, 1, 0, 0, false|coo_warp_kernel<float, 1024, 1, 8, 1, 0, 1, true|coo_segscan_kernel<float, 256, 7, false>|"
    r"cg_update_kernel<double, false>|cg_update_p2p_kernel<double>|cg_direction_p2p_kernel<double>|allgather_push_kernel|csr_vector_kernel<float, 256, 8, 1>")
KEEP = re.compile(r"^(UBLKCP|SYNCS|LDG|STG|LDS|STS|SHFL|BAR|VOTE|MATCH|ATOM|RED|LDGSTS|UTMA|CCTL|MEMBAR|FENCE|ERRBAR|ACQBULK|DEPBAR|NANOSLEEP|LD\.|ST\.)")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, hist, total = None, {}, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void b200sp::", "").replace("b200sp::", "")
            cur = name if PAT.search(name) else None
            if cur:
                hist[cur] = collections.Counter()
                total[cur] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        total[cur] += 1
        op = m.group(1)
        if KEEP.match(op):
            hist[cur][op] += 1
    print("# SASS opcode histogram of the kernels the benchmark runs (cuobjdump -sass cusp_autotuned_b200/libb200sp.so, sm_100a)\n")
    print("UBLKCP = cp.async.bulk (1-D bulk copy through the TMA engine), SYNCS.* = mbarrier operations, LDG.E.128 / .ENL2.256 = 128- /")
    print("256-bit global loads, .EF = evict-first (ld.global.cs), .NA = L1::no_allocate, .CONSTANT = ld.global.nc, ACQBULK / DEPBAR-free")
    print("griddepcontrol shows as the *.ACQBULK / PREEXIT family.\n")
    for k in sorted(hist):
        print(f"## `{k}` — {total[k]} instructions\n")
        print(", ".join(f"`{op}` x {n}" for op, n in sorted(hist[k].items(), key=lambda kv: (-kv[1], kv[0]))) or "(none of the tracked opcodes)")
        print()


if __name__ == "__main__":
    main()
