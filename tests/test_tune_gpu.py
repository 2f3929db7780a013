"""GPU tests of the cusp::ktt replacement (testing/ktt.cu): every configuration of
the space is run and validated; dynamic tuning; cache persistence."""
import os

import numpy as np
import pytest
import torch

import cusp_autotuned_b200 as cusp
from cusp_autotuned_b200 import capi, ktt
from golden import reference_fixtures as G
from helpers import tdev, upload
from oracle import oracle as O

pytestmark = pytest.mark.gpu
OK, LAUNCH_FAILED, VALIDATION_FAILED, UNSUPPORTED = range(4)


def _check_all(results, what):
    """testing/ktt.cu:84-140: no CompilationFailed / ComputationFailed / ValidationFailed"""
    bad = [(r.cfg, r.status, r.max_rel_error) for r in results if r.status in (LAUNCH_FAILED, VALIDATION_FAILED)]
    assert not bad, (what, bad[:3])
    assert sum(r.status == OK for r in results) >= 0.7 * len(results), what


@pytest.mark.parametrize("fmt", ["dia", "ell", "ellr", "csr", "coo", "hyb"])
def test_tune_every_configuration_on_reference_fixtures(fmt, dev, handle):
    """testing/ktt.cu:208-282 (the reference only covers dia/ell/ellr; csr/coo/hyb have no test there)"""
    mats = {k: O.dense_to_coo(v.astype(np.float32)) for k, v in G.MULTIPLY_DENSE.items() if k != "E"}
    for name, coo in mats.items():
        A = O.convert(coo, "ell" if fmt == "ellr" else fmt, **(dict(num_entries_per_row=1) if fmt == "hyb" else {}))
        Ad = upload("ell" if fmt == "ellr" else fmt, A, dev)
        if fmt == "ellr":
            Ad = cusp.ellr_matrix(Ad)
        x = tdev((np.arange(coo["num_cols"]) % 10).astype(np.float32), dev)
        y = torch.zeros(coo["num_rows"], dtype=torch.float32, device=dev)
        # reference computation = the non-tuned path, like ktt.cu:173-181 (ktt::disable + multiply)
        ref = tdev(O.spmv(O.to_ellr(A) if fmt == "ellr" else A, x.cpu().numpy()), dev)
        best, results = ktt.tune(Ad, x, y, reference=ref)
        _check_all(results, (fmt, name))
        assert torch.equal(y, ref)
        ktt.reset_tuning(Ad)


def test_tune_banded_fixtures(dev, handle):
    """testing/ktt.cu:274-281"""
    for rows, cols, step, cnt in G.KTT_BANDED:
        A = O.make_diagonal_symmetric(rows, cols, step, cnt)
        Ad = upload("dia", A, dev)
        x = tdev((np.arange(cols) % 10).astype(np.float32), dev)
        y = torch.zeros(rows, dtype=torch.float32, device=dev)
        best, results = ktt.tune(Ad, x, y, repeats=2)
        _check_all(results, (rows, cols))
        assert np.array_equal(y.cpu().numpy(), O.spmv(A, x.cpu().numpy()))
        assert handle.tune_lookup(Ad.descriptor()) is not None


def test_dynamic_tuning_through_plain_multiply(dev, handle):
    """plain cusp::multiply on ELL/DIA does one tuning step per call while ktt is enabled
    (generic/multiply.inl:141-154); with ktt::disable() it is the fixed default kernel"""
    A = O.poisson(7, (16, 16, 16), np.float32, "dia")
    Ad = upload("dia", A, dev)
    x = tdev(np.random.default_rng(0).uniform(-1, 1, A["num_cols"]).astype(np.float32), dev)
    want = O.spmv(A, x.cpu().numpy())
    ktt.reset_tuning()
    ktt.enable()
    n = len(capi.Handle.cfg_space(capi.FMT_DIA, capi.F32))
    seen = set()
    for i in range(n + 3):
        y = torch.zeros(A["num_rows"], dtype=torch.float32, device=dev)
        before = handle.tune_lookup(Ad.descriptor())
        cusp.multiply(Ad, x, y)
        assert np.array_equal(y.cpu().numpy(), want), i  # every configuration gives the same bits
        r = ktt.multiply(Ad, x, y)
        seen.add((r.cfg.kernel, r.cfg.block_size, r.cfg.unroll, r.cfg.stages, r.cfg.ctas_per_sm))
    assert handle.tune_lookup(Ad.descriptor()) is not None  # space exhausted -> winner cached
    assert len(seen) > 5
    ktt.disable()
    y = torch.zeros(A["num_rows"], dtype=torch.float32, device=dev)
    cusp.multiply(Ad, x, y)
    assert np.array_equal(y.cpu().numpy(), want)
    ktt.enable()
    # explicit configuration (cusp::ktt::multiply(A,x,y,conf))
    y.zero_()
    ktt.multiply(Ad, x, y, capi.Cfg(kernel=capi.K_DIA_LDG, block_size=128, unroll=1))
    assert np.array_equal(y.cpu().numpy(), want)


def test_tuning_cache_save_load_reset(dev, handle, tmp_path):
    A = O.poisson(5, (64, 64), np.float64, "csr")
    Ad = upload("csr", A, dev)
    x = torch.ones(A["num_cols"], dtype=torch.float64, device=dev)
    y = torch.zeros(A["num_rows"], dtype=torch.float64, device=dev)
    best, results = ktt.tune(Ad, x, y, repeats=2)
    assert handle.tune_lookup(Ad.descriptor()).as_dict() == best.as_dict()
    p = str(tmp_path / "tune.txt")
    handle.tune_save(p)
    assert os.path.getsize(p) > 0
    ktt.reset_tuning()
    assert handle.tune_lookup(Ad.descriptor()) is None
    handle.tune_load(p)
    assert handle.tune_lookup(Ad.descriptor()).as_dict() == best.as_dict()
    ktt.reset_tuning(Ad)
    assert handle.tune_lookup(Ad.descriptor()) is None


def test_searcher_order_and_stop_condition_are_honoured_during_the_search(dev, handle):
    """b200sp_tune_ex: the searcher's order decides which configurations run and in which order, the stop condition is
    consulted after every configuration — a budget of 4 runs exactly 4 (cuda/ktt/multiply.h:129-146)"""
    A = O.poisson(5, (96, 80), np.float64, "csr")
    Ad = upload("csr", A, dev)
    d = Ad.descriptor()
    x = tdev(np.random.default_rng(1).uniform(0.5, 1.5, A["num_cols"]), dev)
    y = torch.zeros(A["num_rows"], dtype=torch.float64, device=dev)
    space = capi.Handle.cfg_space(capi.FMT_CSR, capi.F64)
    key = lambda c: tuple(c.as_dict().values())
    order = [40, 3, 77, 12, 5, 90]
    handle.tune_reset(None)
    best, results = handle.tune_ex(d, x, y, order=order, repeats=2)
    assert [key(r.cfg) for r in results] == [key(space[i]) for i in order]
    assert key(best) in {key(space[i]) for i in order}
    assert handle.tune_lookup(d).as_dict() == best.as_dict()
    want = O.spmv(A, x.cpu().numpy())
    assert np.max(np.abs(y.cpu().numpy() - want) / np.abs(want)) <= 1e-12  # y = A x by the winner
    # stop after 4 configurations of the whole space
    seen = []
    handle.tune_reset(None)
    best, results = handle.tune_ex(d, x, y, stop=lambda r: (seen.append(r.milliseconds), len(seen) >= 4)[1], repeats=2)
    assert len(results) == 4 and len(seen) == 4
    assert [key(r.cfg) for r in results] == [key(c) for c in space[:4]]
    handle.tune_reset(None)


def test_tuning_cache_distinguishes_structure(dev, handle):
    """a winner found on a banded CSR is not replayed on a random CSR of the same size class (rows and nnz/row fall in
    the same log2 buckets): the cache key carries the structure class the defaults use (banded / scattered / skewed)"""
    rows, k = 1 << 16, 5
    rng = np.random.default_rng(3)
    banded_cols = np.clip(np.arange(rows)[:, None] + np.arange(-2, 3)[None, :], 0, rows - 1).astype(np.int32)
    random_cols = np.sort(rng.integers(0, rows, (rows, k)), axis=1).astype(np.int32)
    Ap = (np.arange(rows + 1) * k).astype(np.int32)
    mats = {}
    for name, cols in (("banded", banded_cols), ("random", random_cols)):
        A = dict(format="csr", num_rows=rows, num_cols=rows, num_entries=rows * k, row_offsets=Ap,
                 column_indices=cols.reshape(-1), values=np.ones(rows * k, np.float32))
        mats[name] = (A, upload("csr", A, dev))
    x = tdev(rng.integers(-3, 4, rows).astype(np.float32), dev)
    y = torch.zeros(rows, dtype=torch.float32, device=dev)
    handle.tune_reset(None)
    best, _ = handle.tune(mats["banded"][1].descriptor(), x, y, repeats=2)
    assert handle.tune_lookup(mats["banded"][1].descriptor()).as_dict() == best.as_dict()
    assert handle.tune_lookup(mats["random"][1].descriptor()) is None
    for name, (A, Ad) in mats.items():  # products stay exact with and without a cached winner
        cusp.multiply(Ad, x, y)
        assert np.array_equal(y.cpu().numpy(), O.spmv(A, x.cpu().numpy())), name
    handle.tune_reset(None)
