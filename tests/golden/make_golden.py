"""Generates tests/golden/ref_spmv.npz by running the REFERENCE's own host loops
(oracle/_ref/libcuspref.so, built from /root/reference by oracle/Makefile) on
seeded inputs.  Run in the build container (needs /root/reference); the GPU box
only reads the committed .npz.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def cases():
    rng = np.random.default_rng(20261018)
    out = {}
    # stencil operators with non-integer x (tolerance / bit parity), fp32 + fp64
    for name, st, grid in (("p5", 5, (13, 9)), ("p7", 7, (7, 6, 5)), ("p9", 9, (6, 7)), ("p27", 27, (4, 3, 5))):
        for dt in (np.float32, np.float64):
            dia = O.poisson(st, grid, dt, "dia")
            x = rng.uniform(0.5, 1.5, dia["num_cols"]).astype(dt)
            y0 = rng.uniform(-1, 1, dia["num_rows"]).astype(dt)
            for fmt in ("csr", "coo", "dia", "ell", "hyb"):
                A = O.convert(dia, fmt)
                key = f"{name}_{np.dtype(dt).name}_{fmt}"
                out[key + "_x"] = x
                out[key + "_y0"] = y0
                out[key + "_y"] = O.spmv(A, x, impl="ref")
                out[key + "_yacc"] = O.spmv(A, x, y0, accumulate=True, impl="ref")
    # gallery::random (ragged rows, duplicates removed), values made non-trivial
    for m, n, s in ((24, 24, 150), (24, 12, 20), (300, 257, 4000)):
        for dt in (np.float32, np.float64):
            coo = O.gallery_random(m, n, s, dt, "coo")
            coo["values"] = rng.uniform(0.5, 1.5, coo["num_entries"]).astype(dt)
            x = rng.uniform(0.5, 1.5, n).astype(dt)
            for fmt in ("csr", "coo", "ell", "hyb"):
                A = O.convert(coo, fmt)
                key = f"rand{m}x{n}_{np.dtype(dt).name}_{fmt}"
                out[key + "_x"] = x
                out[key + "_vals"] = coo["values"]
                out[key + "_y"] = O.spmv(A, x, impl="ref")
    return out


if __name__ == "__main__":
    assert O.ref_available(), "oracle/_ref/libcuspref.so missing: run `make -C oracle` where /root/reference exists"
    data = cases()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_spmv.npz")
    np.savez_compressed(path, **data)
    print(f"wrote {path}: {len(data)} arrays, {os.path.getsize(path)} bytes")
