"""Static check on libb200sp.so's SASS: no ring-stage release (consumer-side SYNCS.ARRIVE) may be
issued while a register loaded from shared memory (LDS) in the preceding straight-line window is
still unconsumed.  On sm_100a the arrive is not ordered behind outstanding LDS (common.cuh,
consume_before_release), so such a schedule lets the producer's next bulk copy overwrite a stage
that is still being read.

  python tools/check_release_order.py [path/to/libb200sp.so]   -> exit status 1 if a site is flagged
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = ("dia_bulk", "ell_bulk", "csr_ring", "coo_ring", "csr_spmm_ring")
WINDOW = 300  # instructions scanned before each arrive


def scan(lib):
    return scan_text(subprocess.check_output(["cuobjdump", "-sass", lib], text=True))


def scan_text(txt):
    """(release sites, flagged sites) of a `cuobjdump -sass` listing"""
    sites, flagged = 0, []
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0]
        if not any(k in name for k in KERNELS):
            continue
        ops = []
        for line in f.split("\n"):
            m = re.search(r"/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?(\S+)\s*(.*?);", line)
            if m:
                ops.append((int(m.group(1), 16), m.group(3), m.group(4)))
        for i, (pc, op, _) in enumerate(ops):
            if not (op.startswith("SYNCS.ARRIVE") and "A1T0" in op):  # consumer release (no expect_tx)
                continue
            sites += 1
            pending = {}
            for _, opj, aj in ops[max(0, i - WINDOW):i]:
                dst = aj.split(",")[0].strip()
                if opj.startswith("LDS"):
                    width = 2 if ".64" in opj else (4 if ".128" in opj else 1)
                    if re.fullmatch(r"R\d+", dst):
                        for k in range(width):
                            pending["R%d" % (int(dst[1:]) + k)] = True
                    continue
                all_operands = opj.startswith(("ST", "ISETP", "FSETP", "DSETP", "LDG", "BRA", "SHFL"))
                rest = aj if all_operands else (aj.split(",", 1)[1] if "," in aj else "")
                wide = opj.startswith(("DADD", "DMUL", "DFMA", "DSETP")) or ".64" in aj
                for reg in re.findall(r"\bR(\d+)\b", rest):
                    pending.pop("R" + reg, None)
                    if wide:
                        pending.pop("R%d" % (int(reg) + 1), None)
                if not opj.startswith("ST"):
                    pending.pop(dst, None)  # overwritten
            if pending:
                flagged.append((name, hex(pc), sorted(pending)))
    return sites, flagged


if __name__ == "__main__":
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cusp_autotuned_b200", "libb200sp.so")
    sites, flagged = scan(lib)
    for f in flagged:
        print("unconsumed LDS before release:", *f)
    print(f"{sites} release sites, {len(flagged)} flagged")
    sys.exit(1 if flagged or sites == 0 else 0)
