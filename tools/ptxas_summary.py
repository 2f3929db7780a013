"""Compile one .cu of csrc/ with -Xptxas -v and print one line per kernel: registers, spills, smem.
usage: python tools/ptxas_summary.py spmv_coo_warp.cu [filter-regex]"""
import re, subprocess, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csrc = os.path.join(root, "cusp_autotuned_b200", "csrc")
src = sys.argv[1]
flt = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
       "-I../../include", "-I.", "-Xcompiler", "-fPIC,-fvisibility=default", "-Xptxas", "-v", "-c", src, "-o", "/dev/null"]
out = subprocess.run(cmd, cwd=csrc, capture_output=True, text=True).stderr
name = None
rows = {}
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void b200sp::", "")
        rows[name] = dict(regs=0, spill=0, smem=0)
        continue
    m = re.search(r"(\d+) bytes spill stores", line)
    if m and name:
        rows[name]["spill"] = int(m.group(1))
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        rows[name]["regs"] = int(m.group(1))
        s = re.search(r"(\d+) bytes smem", line)
        rows[name]["smem"] = int(s.group(1)) if s else 0
    if "error" in line:
        print(line)
for k, v in sorted(rows.items()):
    if flt and not flt.search(k):
        continue
    print(f"{v['regs']:4d} regs {v['spill']:5d} spill {v['smem']:6d} smem  {k}")
