"""cusp::convert on the device (csrc/convert.cu): layouts bit-identical to the
reference's rules — against the reference's own golden arrays (testing/convert.cu:63-200,
405-497), the oracle's restatement on random matrices, the device gallery builders at
BASELINE-like sizes, and closed forms where the three-level scan is exercised."""
import numpy as np
import pytest
import torch

from cusp_autotuned_b200 import capi, convert, gallery
from cusp_autotuned_b200.matrix import csr_matrix
from golden import reference_fixtures as F
from helpers import upload
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _eq(t, a):
    return np.array_equal(t.cpu().numpy(), a)


def test_reference_golden_layouts_alignment_1(dev):
    """the 4x4 / 7-entry example of testing/convert.cu in every target format"""
    csr = upload("csr", F.CONVERT_CSR, dev)
    coo = convert.csr_to_coo(csr)
    assert _eq(coo.row_indices, F.CONVERT_COO["row_indices"])
    back = convert.coo_to_csr(coo)
    assert _eq(back.row_offsets, F.CONVERT_CSR["row_offsets"])
    ell = convert.csr_to_ell(csr, alignment=1)
    assert ell.num_cols_per_row == 3 and ell.pitch == 4 and ell.num_entries == 7
    assert _eq(ell.column_indices, F.CONVERT_ELL["column_indices"]) and _eq(ell.values, F.CONVERT_ELL["values"])
    dia = convert.csr_to_dia(csr, alignment=1)
    assert dia.pitch == 4 and _eq(dia.diagonal_offsets, F.CONVERT_DIA["diagonal_offsets"])
    assert _eq(dia.values, F.CONVERT_DIA["values"])
    hyb = convert.csr_to_hyb(csr, num_entries_per_row=1, alignment=1)
    assert _eq(hyb.ell.column_indices, F.CONVERT_HYB["ell"]["column_indices"])
    assert _eq(hyb.ell.values, F.CONVERT_HYB["ell"]["values"])
    assert _eq(hyb.coo.row_indices, F.CONVERT_HYB["coo"]["row_indices"])
    assert _eq(hyb.coo.column_indices, F.CONVERT_HYB["coo"]["column_indices"])
    assert _eq(hyb.coo.values, F.CONVERT_HYB["coo"]["values"])


@pytest.mark.parametrize("ndt", (np.float32, np.float64))
def test_random_matrix_against_the_oracle(ndt, dev, handle):
    rng = np.random.default_rng(21)
    coo = O.gallery_random(5000, 4200, 60000, ndt, "coo")  # ragged rows, some empty
    coo["values"] = rng.uniform(0.5, 1.5, coo["num_entries"]).astype(ndt)
    coo["values"][::97] = 0  # stored zeros: ELL num_entries = nnz - zeros (csr_to_other.h:205-212)
    csr = O.convert(coo, "csr")
    csr_d = convert.coo_to_csr(upload("coo", coo, dev))
    assert _eq(csr_d.row_offsets, csr["row_offsets"])
    assert _eq(convert.csr_to_coo(csr_d).row_indices, coo["row_indices"])
    info = handle.csr_convert_query(csr_d.num_rows, csr_d.num_cols, csr_d.num_entries, csr_d.row_offsets,
                                    csr_d.column_indices)
    lens = np.diff(csr["row_offsets"])
    assert info.max_entries_per_row == lens.max()
    assert info.hyb_entries_per_row == O.optimal_entries_per_row(csr["row_offsets"])
    assert info.hyb_coo_entries == np.maximum(lens - info.hyb_entries_per_row, 0).sum()
    assert handle.count_zeros(csr_d.values) == int((coo["values"] == 0).sum())
    for K in (0, 1, 7, int(lens.max())):
        if K == 0:
            want, got = O.convert(csr, "ell"), convert.csr_to_ell(csr_d)
        else:
            want, got = O.convert(csr, "ell", num_entries_per_row=K), convert.csr_to_ell(csr_d, K)
        assert got.pitch == want["pitch"] and got.num_cols_per_row == want["num_cols_per_row"]
        assert _eq(got.column_indices, want["column_indices"]) and _eq(got.values, want["values"]), K
    assert convert.csr_to_ell(csr_d).num_entries == O.convert(csr, "ell")["num_entries"]
    for K in (None, 2, 5):
        want = O.convert(csr, "hyb") if K is None else O.convert(csr, "hyb", num_entries_per_row=K)
        got = convert.csr_to_hyb(csr_d, K)
        assert _eq(got.ell.column_indices, want["ell"]["column_indices"]) and _eq(got.ell.values, want["ell"]["values"])
        assert _eq(got.coo.row_indices, want["coo"]["row_indices"])
        assert _eq(got.coo.column_indices, want["coo"]["column_indices"])
        assert _eq(got.coo.values, want["coo"]["values"]), K


@pytest.mark.parametrize("ndt", (np.float32, np.float64))
def test_csr_to_dia_against_the_oracle_and_fill_guard(ndt, dev):
    A = O.poisson(7, (13, 11, 9), ndt, "csr")
    A["values"] = (A["values"] * np.random.default_rng(2).uniform(0.5, 1.5, len(A["values"]))).astype(ndt)
    want = O.convert(A, "dia")
    got = convert.csr_to_dia(upload("csr", A, dev))
    assert got.pitch == want["pitch"] and got.num_diagonals == len(want["diagonal_offsets"])
    assert _eq(got.diagonal_offsets, want["diagonal_offsets"]) and _eq(got.values, want["values"])
    # rectangular, diagonals on both sides
    B = O.gallery_random(300, 500, 900, ndt, "csr")
    want = O.convert(B, "dia")
    got = convert.csr_to_dia(upload("csr", B, dev))
    assert _eq(got.diagonal_offsets, want["diagonal_offsets"]) and _eq(got.values, want["values"])
    # a random pattern at scale is refused like in the reference (> 3x fill-in on > 1e6 slots)
    C = O.gallery_random(4000, 4000, 20000, ndt, "csr")
    with pytest.raises(capi.B200spError):
        convert.csr_to_dia(upload("csr", C, dev))


def test_full_size_conversions_match_the_device_gallery(dev):
    """poisson7pt 160^3 (4.1 M rows): CSR -> DIA / ELL reproduce the gallery's arrays; CSR <-> COO round trip"""
    n = 160
    csr = gallery.poisson("csr", 7, (n, n, n), dtype=torch.float64)
    dia = convert.csr_to_dia(csr)
    ref = gallery.poisson("dia", 7, (n, n, n), dtype=torch.float64)
    assert dia.pitch == ref.pitch and torch.equal(dia.diagonal_offsets, ref.diagonal_offsets)
    assert torch.equal(dia.values, ref.values)
    ell = convert.csr_to_ell(csr)
    ref = gallery.poisson("ell", 7, (n, n, n), dtype=torch.float64)
    assert ell.pitch == ref.pitch and ell.num_cols_per_row == 7 and ell.num_entries == ref.num_entries
    assert torch.equal(ell.column_indices, ref.column_indices) and torch.equal(ell.values, ref.values)
    coo = convert.csr_to_coo(csr)
    assert bool((coo.row_indices[1:] >= coo.row_indices[:-1]).all())
    assert torch.equal(convert.coo_to_csr(coo).row_offsets, csr.row_offsets)


def test_tail_extraction_over_twenty_million_rows(dev, handle):
    """three scan levels (> 4096^2 rows): row r has r % 3 entries, K = 1 keeps one, the tail is closed-form"""
    rows = 20_000_003
    lens = (torch.arange(rows, device=dev, dtype=torch.int64) % 3)
    Ap = torch.zeros(rows + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=Ap[1:])
    nnz = int(Ap[-1].item())
    Ap = Ap.to(torch.int32)
    Aj = (torch.arange(nnz, device=dev, dtype=torch.int64) % 1000).to(torch.int32)
    Ax = torch.arange(nnz, device=dev, dtype=torch.float32)
    A = csr_matrix(rows, 1000, Ap, Aj, Ax)
    info = handle.csr_convert_query(rows, 1000, nnz, Ap)
    assert info.max_entries_per_row == 2
    hyb = convert.csr_to_hyb(A, num_entries_per_row=1)
    tail_rows = torch.nonzero(lens == 2).flatten()
    assert hyb.coo.num_entries == tail_rows.numel()
    assert torch.equal(hyb.coo.row_indices.to(torch.int64), tail_rows)
    src = Ap.to(torch.int64)[tail_rows] + 1  # second entry of every 2-entry row
    assert torch.equal(hyb.coo.values, Ax[src]) and torch.equal(hyb.coo.column_indices, Aj[src])
    first = torch.where(lens > 0, Ax[torch.clamp(Ap[:-1].to(torch.int64), max=nnz - 1)], torch.zeros((), device=dev))
    assert torch.equal(hyb.ell.values[:rows], first)


def _same_layout(D, H):
    """device container D against the oracle's dict H: every array bit for bit"""
    f = H["format"]
    if f == "csr":
        return _eq(D.row_offsets, H["row_offsets"]) and _eq(D.column_indices, H["column_indices"]) and _eq(D.values, H["values"])
    if f == "coo":
        return _eq(D.row_indices, H["row_indices"]) and _eq(D.column_indices, H["column_indices"]) and _eq(D.values, H["values"])
    if f == "ell":
        return (D.num_cols_per_row == H["num_cols_per_row"] and D.pitch == H["pitch"] and
                _eq(D.column_indices, H["column_indices"]) and _eq(D.values, H["values"]))
    if f == "dia":
        return D.pitch == H["pitch"] and _eq(D.diagonal_offsets, H["diagonal_offsets"]) and _eq(D.values, H["values"])
    if f == "hyb":
        return _same_layout(D.ell, H["ell"]) and _same_layout(D.coo, H["coo"])
    raise ValueError(f)


@pytest.mark.parametrize("ndt", (np.float32, np.float64))
def test_every_source_times_destination_on_the_device(ndt, dev, handle):
    """cusp::convert for all 5 x 5 format pairs with device containers never leaves the device (convert.convert ->
    b200sp_{dia,ell,hyb}_to_csr_*, b200sp_dia_to_ell, b200sp_csr_to_*): every array of the result equals the oracle's
    restatement of generic/conversions/{dia,ell,hyb,csr,coo}_to_other.h — stencil operators (every format incl. DIA),
    a ragged random matrix with stored zeros and empty rows (no DIA), and a HYB whose split leaves a COO tail."""
    rng = np.random.default_rng(41)
    cases = [("stencil5", O.poisson(5, (31, 17), ndt, "coo"), True), ("stencil7", O.poisson(7, (9, 8, 7), ndt, "coo"), True)]
    rnd = O.gallery_random(3000, 2600, 40000, ndt, "coo")
    rnd["values"] = rng.uniform(0.5, 1.5, rnd["num_entries"]).astype(ndt)
    rnd["values"][::53] = 0
    cases.append(("random", rnd, False))
    for name, coo, banded in cases:
        fmts = ("csr", "coo", "ell", "hyb", "dia") if banded else ("csr", "coo", "ell", "hyb")
        for src in fmts:
            kw = dict(num_entries_per_row=2) if src == "hyb" else {}
            Sh = O.convert(coo, src, **kw)
            Sd = upload(src, Sh, dev)
            for dst in fmts:
                if dst == src:
                    continue
                want = O.convert(Sh, dst)
                l0 = handle.launch_count
                got = convert.convert(Sd, dst)
                assert handle.launch_count > l0 or dst == "hyb", (name, src, dst)  # device kernels did the work
                assert _same_layout(got, want), (name, src, dst)
    assert not hasattr(convert, "to_host")  # no host detour in the module


@pytest.mark.parametrize("tdt", (torch.float32, torch.float64))
def test_conversions_at_256_cubed_stay_on_the_device(tdt, dev, handle):
    """poisson7pt 256^3 (BASELINE configs[1]): DIA -> CSR / COO / ELL and ELL -> CSR, HYB -> CSR on the device equal
    the arrays the device gallery builds directly (themselves bit-identical to the oracle pipeline,
    tests/test_gallery_gpu.py) — 117M entries per conversion, no PCIe traffic."""
    n = 256
    dia = gallery.poisson("dia", 7, (n, n, n), dtype=tdt)
    csr_ref = gallery.poisson("csr", 7, (n, n, n), dtype=tdt)
    csr = convert.dia_to_csr(dia)
    assert csr.num_entries == 117047296
    assert torch.equal(csr.row_offsets, csr_ref.row_offsets) and torch.equal(csr.column_indices, csr_ref.column_indices)
    assert torch.equal(csr.values, csr_ref.values)
    del csr
    ell = convert.dia_to_ell(dia)
    ell_ref = gallery.poisson("ell", 7, (n, n, n), dtype=tdt)
    assert ell.pitch == ell_ref.pitch and ell.num_cols_per_row == 7
    assert torch.equal(ell.column_indices, ell_ref.column_indices) and torch.equal(ell.values, ell_ref.values)
    del dia, ell_ref
    back = convert.ell_to_csr(ell)
    assert torch.equal(back.row_offsets, csr_ref.row_offsets) and torch.equal(back.column_indices, csr_ref.column_indices)
    assert torch.equal(back.values, csr_ref.values)
    del back, ell
    hyb = convert.csr_to_hyb(csr_ref, num_entries_per_row=6)
    back = convert.hyb_to_csr(hyb)
    assert torch.equal(back.row_offsets, csr_ref.row_offsets) and torch.equal(back.column_indices, csr_ref.column_indices)
    assert torch.equal(back.values, csr_ref.values)
