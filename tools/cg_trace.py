"""Summarise B200SP_CG_TRACE files (<path>.<rank>, written by b200sp_cg_dist on the peer-memory path): where an
iteration's time goes on every rank — K2 poll wait (all ranks' <y,p>), K2 body, K2->K3 gap, K3 poll wait (all ranks'
<r,r>), K3 body, and K3 end -> next K2 entry (= the SpMV K1 plus its launch gap and halo wait).

  python tools/cg_trace.py <path> [skip_first=3]"""
import glob
import json
import sys

import numpy as np


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    out = {}
    for f in sorted(glob.glob(path + ".*")):
        rank = f.rsplit(".", 1)[1]
        if not rank.isdigit():
            continue
        rows = [[int(v) for v in ln.split()] for ln in open(f) if ln.strip() and not ln.startswith("#")]
        a = np.array(rows, dtype=np.int64)
        a = a[(a[:, 1:] > 0).all(axis=1)][skip:]
        if len(a) < 2:
            continue
        k2e, k2p, k2x, k3e, k3p, k3x = (a[:, i].astype(np.float64) for i in range(1, 7))
        us = lambda v: round(float(np.mean(v)) / 1e3, 2)
        out[rank] = {"iters": int(len(a)), "k2_poll_wait_us": us(k2p - k2e), "k2_body_us": us(k2x - k2p), "k2_to_k3_gap_us": us(k3e - k2x),
                     "k3_poll_wait_us": us(k3p - k3e), "k3_body_us": us(k3x - k3p), "k3_end_to_next_k2_us": us(k2e[1:] - k3x[:-1]),
                     "iteration_us": us(k2e[1:] - k2e[:-1])}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
