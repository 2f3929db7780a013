"""Per-launch summary of an .ncu-rep (ncu -i <rep> --page raw --csv): duration, DRAM bytes, unit throughputs, stalls.
usage: python tools/ncu_summary.py <file.ncu-rep> [--md]"""
import csv
import io
import re
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "MB rd", None), ("dram__bytes_write.sum", "MB wr", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1), ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex%", 1),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%", 1), ("l1tex__t_sector_hit_rate.pct", "L1hit%", 1),
        ("lts__t_sector_hit_rate.pct", "L2hit%", 1), ("launch__registers_per_thread", "regs", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
        ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "st_long_sb", 1),
        ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "st_mio", 1),
        ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "st_bar", 1),
        ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "st_lg", 1),
        ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "st_short_sb", 1)]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {k: hdr.index(k) for k, _, _ in WANT if k in hdr}
    kn = hdr.index("Kernel Name")
    cols = [(k, lab, sc) for k, lab, sc in WANT if k in idx]
    print("| kernel | " + " | ".join(lab for _, lab, _ in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[kn]).replace("void b200sp::", "").replace("b200sp::", "")
        vals = []
        for k, lab, sc in cols:
            v = r[idx[k]].replace(",", "")
            try:
                f = float(v)
                u = units[idx[k]]
                if sc is None:  # bytes -> MB according to the unit ncu chose
                    f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                elif lab == "us":
                    f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1e-3)
                vals.append(f"{f:.1f}")
            except ValueError:
                vals.append(v)
        print(f"| {name[:64]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
