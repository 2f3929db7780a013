"""CPU suite: the C-ABI library loads and exports every symbol include/b200sp.h
declares; no compute calls are made (no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from cusp_autotuned_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200sp.h")


def _declared_symbols():
    """expand the header with the C preprocessor and collect every b200sp_* function"""
    out = subprocess.check_output(["gcc", "-E", "-P", HEADER], text=True)
    names = set(re.findall(r"\b(b200sp_[a-z0-9_]+)\s*\(", out))
    return sorted(names)


def test_header_is_plain_c():
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", HEADER])
    txt = open(HEADER).read()
    assert "torch" not in txt and "at::Tensor" not in txt  # plain pointers and sizes only


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 50
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(capi.EXPORTED_SYMBOLS) == declared  # the binding covers the whole ABI


def test_every_declaration_cites_the_reference():
    txt = open(HEADER).read()
    for needle in ("csr_vector_spmv.h", "ell_spmv.h", "dia_spmv.h", "coo_flat_spmv.h", "generic/multiply/spmv.h",
                   "cusp/krylov/detail/cg.inl", "cusp/detail/monitor.inl", "generic/blas.h", "ktt.inl",
                   "stencil.inl", "ellr_matrix"):
        assert needle in txt, needle


def test_version_and_strings():
    lib = capi.load_library()
    assert lib.b200sp_version() == 100
    assert lib.b200sp_status_string(0) == b"ok"
    assert lib.b200sp_status_string(1) == b"invalid input"


def test_cfg_spaces_enumerate():
    sizes = {f: len(capi.Handle.cfg_space(f, capi.F32)) for f in range(6)}
    assert sizes[capi.FMT_CSR] == 62 + 45 + 4 and sizes[capi.FMT_COO] == 10 + 48 + 6  # segscan + ring (8 shapes x 2 stages x 3 ctas/sm) + warp (2 widths x 3 unrolls)
    assert sizes[capi.FMT_ELL] == sizes[capi.FMT_DIA] == sizes[capi.FMT_ELLR] == 63
    for c in capi.Handle.cfg_space(capi.FMT_CSR, capi.F64):
        assert c.kernel in (capi.K_CSR_VECTOR, capi.K_CSR_STREAM, capi.K_CSR_RING, capi.K_CSR_BALANCED)
        assert c.block_size in (128, 256, 512)
        if c.kernel == capi.K_CSR_VECTOR:
            assert c.threads_per_row in (1, 2, 4, 8, 16, 32) and c.unroll in (1, 2, 4)
        elif c.kernel == capi.K_CSR_BALANCED:
            assert c.unroll in (5, 7, 9)
        elif c.kernel == capi.K_CSR_RING:
            assert c.unroll in (4, 8, 16) and c.stages in (2, 3, 4) and c.ctas_per_sm in (2, 4, 6)
        else:
            assert c.unroll in (4, 8, 16)


def test_poisson_entry_count_closed_form():
    n = capi.poisson_num_entries
    assert n(5, 512, 512, 1, 0, 512 * 512) == 1308672
    assert n(7, 256, 256, 256, 0, 256 ** 3) == 117047296    # config 2
    assert n(7, 512, 512, 512, 0, 512 ** 3) == 937951232    # config 5
    assert n(7, 2, 2, 2, 0, 8) == 32 and n(5, 2, 3, 1, 0, 6) == 20
    # additive over row blocks
    assert sum(n(7, 6, 5, 8, r0, 60) for r0 in range(0, 240, 60)) == n(7, 6, 5, 8, 0, 240)


def test_no_cpu_fallback_without_gpu():
    """the product path fails loudly when there is no device"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.B200spError) as e:
        capi.Handle()
    assert "no CPU fallback" in str(e.value)
    import cusp_autotuned_b200 as cusp
    with pytest.raises(capi.B200spError):
        cusp.default_handle()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing in the package or the C++ headers refers to it"""
    bad = []
    for base in ("cusp_autotuned_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".inl")):
                    s = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|liboracle|libcuspref|oracle/", s):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_only_sm100a_code_in_the_library():
    if not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "-lelf", capi.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_ring_stages_are_released_after_their_loads_are_consumed():
    """SASS check (tools/check_release_order.py): no consumer-side mbarrier arrive with an
    unconsumed shared-memory load in front of it — the hardware does not order the two"""
    if not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import check_release_order
    sites, flagged = check_release_order.scan(capi.LIB_PATH)
    assert sites >= 40 and not flagged, flagged


def test_release_order_checker_flags_an_unconsumed_load():
    """the checker itself: an arrive right after an LDS whose register nobody read is flagged, the same
    listing with a consumer in between is not"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import check_release_order
    head = "\n\tFunction : _ZN6b200sp15coo_ring_kernelIfLi256ELi7EEEvNS_7CooArgsIT_EEix\n"
    lds = "        /*0010*/                   LDS R4, [R2] ;\n        /*0020*/                   LDS.64 R6, [R2+0x8] ;\n"
    use = "        /*0030*/                   FADD R9, R4, R6 ;\n        /*0040*/                   FADD R9, R9, R7 ;\n"
    arrive = "        /*0050*/              @!P2 SYNCS.ARRIVE.TRANS64.A1T0 RZ, [R18+URZ], RZ ;\n"
    producer = "        /*0060*/                   SYNCS.ARRIVE.TRANS64 RZ, [UR9], R16 ;\n"  # expect_tx arrive: not a release
    sites, flagged = check_release_order.scan_text(head + lds + arrive + producer)
    assert sites == 1 and len(flagged) == 1 and flagged[0][2] == ["R4", "R6", "R7"]
    sites, flagged = check_release_order.scan_text(head + lds + use + arrive + producer)
    assert sites == 1 and not flagged
    other = "\n\tFunction : _ZN6b200sp11dot_kernelIfEEvv\n" + lds + arrive  # not a ring kernel: ignored
    assert check_release_order.scan_text(other) == (0, [])


def test_containers_refuse_wrong_index_dtype_and_lengths():
    """torch's default index dtype is int64; the ABI takes int32 and would reinterpret the bytes.  The containers
    refuse it, and any array whose length contradicts the shape, before a pointer reaches the library."""
    import torch
    from cusp_autotuned_b200 import matrix as M
    i64 = torch.arange(4)
    i32 = i64.to(torch.int32)
    v3 = torch.ones(3)
    for bad in (lambda: M.csr_matrix(3, 3, i64, i32[:3], v3),                      # int64 offsets
                lambda: M.csr_matrix(3, 3, i32, i64[:3], v3),                      # int64 columns
                lambda: M.csr_matrix(4, 3, i32, i32[:3], v3),                      # offsets length != rows + 1
                lambda: M.csr_matrix(3, 3, i32, i32[:2], v3),                      # columns length != nnz
                lambda: M.coo_matrix(3, 3, i64[:3], i32[:3], v3),
                lambda: M.coo_matrix(3, 3, i32[:2], i32[:3], v3),
                lambda: M.ell_matrix(3, 3, 3, 2, 4, i32, torch.ones(8)),            # indices shorter than K * pitch
                lambda: M.ell_matrix(3, 3, 3, 1, 2, i32, torch.ones(8)),            # pitch < rows
                lambda: M.dia_matrix(3, 3, 3, i64[:1], 3, v3),
                lambda: M.dia_matrix(3, 3, 3, i32[:2], 3, v3),                      # values shorter than ndiag * pitch
                lambda: M.csr_matrix(3, 3, i32, i32[:3], torch.ones(3, dtype=torch.float16))):
        with pytest.raises(capi.InvalidInput):
            bad()
    # well-formed host tensors get as far as the device check ("no CPU fallback" behind device containers)
    with pytest.raises(capi.InvalidInput) as e:
        M.csr_matrix(3, 3, i32, i32[:3], v3)
    assert "host tensor" in str(e.value)
