"""The templated C++ host layer (include/cusp/*.h) over the C ABI.

tests/cpp/*.cpp restate the reference's own unit tests (testing/multiply.cu,
blas.cu, cg.cu, convert.cu, poisson.cu, format_utils.cu, ktt.cu and
examples/Views/cg_raw.cu) against the drop-in headers; one binary,
tests/cpp/build/cusp_api_tests, runs them.  CPU suite: the host_memory half
(pure host loops — the reference's host path) and that a translation unit using
only the public headers compiles.  GPU suite: the whole binary; every
device_memory test goes through libb200sp.so.
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
BIN = os.path.join(CPP, "build", "cusp_api_tests")


def _build():
    subprocess.check_call(["make", "-s", "-j", "8", "-C", os.path.join(ROOT, "cusp_autotuned_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-j", "8", "-C", CPP])
    assert os.path.exists(BIN)


def _run(*args, timeout=900):
    _build()
    p = subprocess.run([BIN, *args], capture_output=True, text=True, timeout=timeout)
    return p.returncode, p.stdout + p.stderr


def _summary(out):
    m = re.search(r"SUMMARY passed=(\d+) failed=(\d+) skipped=(\d+)", out)
    assert m, out[-2000:]
    return tuple(int(g) for g in m.groups())


def test_cpp_host_memory_suite():
    rc, out = _run("--host-only")
    passed, failed, skipped = _summary(out)
    assert rc == 0 and failed == 0, out[-4000:]
    assert passed >= 40 and skipped >= 40  # the device half waits for the GPU suite


def test_cpp_lists_reference_test_names():
    rc, out = _run("--list")
    assert rc == 0
    for name in ("TestSparseMatrixVectorMultiply<csr,float,device>", "TestScaledSparseMatrixVectorMultiply<hyb,double,device>",
                 "TestConjugateGradient<device_memory>", "TestKttBanded<Dia>", "TestAxpby<device_memory>",
                 "TestCgRawPointers", "TestConvertAcrossSpaces", "TestGeneralizedMinRes<device_memory>",
                 "TestReadMatrixMarketFileToCsrMatrix<device_memory>"):
        assert name in out, name


def test_public_headers_are_self_contained(tmp_path):
    """each public header compiles on its own with the host compiler (no Thrust, no nvcc needed)"""
    headers = ["cusp/array1d.h", "cusp/array2d.h", "cusp/coo_matrix.h", "cusp/csr_matrix.h", "cusp/dia_matrix.h",
               "cusp/ell_matrix.h", "cusp/hyb_matrix.h", "cusp/convert.h", "cusp/copy.h", "cusp/format_utils.h",
               "cusp/multiply.h", "cusp/blas/blas.h", "cusp/blas.h", "cusp/monitor.h", "cusp/krylov/cg.h",
               "cusp/linear_operator.h", "cusp/gallery/poisson.h", "cusp/gallery/random.h", "cusp/ktt/ktt.h",
               "cusp/ktt/ellr_matrix.h", "cusp/ktt/matrix_generation.h", "cusp/functional.h", "cusp/exception.h",
               "cusp/krylov/bicgstab.h", "cusp/krylov/cr.h", "cusp/krylov/gmres.h", "cusp/precond/diagonal.h",
               "cusp/io/matrix_market.h", "cusp/print.h", "cusp/complex.h"]
    for h in headers:
        src = tmp_path / "tu.cpp"
        src.write_text(f"#include <{h}>\nint main() {{ return 0; }}\n")
        subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                               "-I", "/usr/local/cuda/include", str(src)])


@pytest.mark.gpu
def test_cpp_full_suite_on_gpu():
    rc, out = _run()
    passed, failed, skipped = _summary(out)
    assert rc == 0 and failed == 0 and skipped == 0, out[-6000:]
    assert passed >= 100


REFERENCE_EXAMPLES = [  # the reference's own example programs that touch the hot path, compiled UNMODIFIED
    "examples/Views/array1d.cu", "examples/Views/array2d_raw.cu", "examples/Views/cg_raw.cu", "examples/Views/csr_raw.cu",
    "examples/Views/csr_view.cu", "examples/Solvers/cg.cu", "examples/Solvers/bicgstab.cu", "examples/Solvers/cr.cu",
    "examples/MatrixFormats/coo.cu", "examples/MatrixFormats/csr.cu", "examples/MatrixFormats/dia.cu",
    "examples/MatrixFormats/ell.cu", "examples/MatrixFormats/hyb.cu", "examples/Gallery/poisson.cu",
    "examples/InputOutput/matrix_market.cu", "examples/Preconditioners/diagonal.cu",
    "examples/Solvers/gmres.cu", "examples/LinearOperator/stencil.cu",  # thrust::fill / raw_pointer_cast on the containers
]


def test_reference_examples_compile_unmodified_against_the_drop_in_headers(tmp_path):
    """drop-in evidence: the reference's example sources (read where they lie under /root/reference, nothing copied)
    compile with nvcc for sm_100a against include/cusp — Thrust's device_ptr, cusp::print, views over raw cudaMalloc
    pointers, Thrust algorithms and thrust::raw_pointer_cast on the containers' iterators (under nvcc they ARE
    thrust::device_ptr), the solvers, MatrixMarket I/O included.  Not covered: Gallery/diffusion.cu (off-path header)."""
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "examples")):
        pytest.skip("reference sources not present (GPU box)")
    import shutil
    from concurrent.futures import ThreadPoolExecutor
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not available")

    def compile_one(rel):
        obj = tmp_path / (rel.replace("/", "_") + ".o")
        p = subprocess.run(["nvcc", "-std=c++17", "-x", "cu", "-w", "-c", "-gencode", "arch=compute_100a,code=sm_100a",
                            "-I", os.path.join(ROOT, "include"), "-o", str(obj), os.path.join(ref, rel)],
                           capture_output=True, text=True, timeout=600)
        return rel, p.returncode, p.stderr[-600:]
    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, REFERENCE_EXAMPLES))
    bad = [(r, err) for r, rc, err in results if rc != 0]
    assert not bad, bad


def test_cpp_suite_builds_with_nvcc_thrust_iterators(tmp_path):
    """under nvcc the device containers iterate with thrust::device_ptr / thrust::device_reference (cusp/array1d.h):
    the restated test programs must compile in that mode too (every device instantiation), and their host_memory half,
    linked from the nvcc objects, must pass"""
    import shutil
    from concurrent.futures import ThreadPoolExecutor
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not available")
    srcs = ["main.cpp", "test_views.cpp", "test_containers.cpp", "test_multiply.cpp", "test_krylov.cpp", "test_blas.cpp",
            "test_dispatch.cpp"]

    def compile_one(name):
        obj = tmp_path / (name + ".o")
        p = subprocess.run(["nvcc", "-std=c++17", "-x", "cu", "-w", "-c", "-fmad=false", "-Xcompiler", "-ffp-contract=off",
                            "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"), "-I", CPP,
                            "-o", str(obj), os.path.join(CPP, name)], capture_output=True, text=True, timeout=900)
        return name, p.returncode, p.stderr[-800:], str(obj)
    with ThreadPoolExecutor(max_workers=7) as ex:
        results = list(ex.map(compile_one, srcs))
    bad = [(n, err) for n, rc, err, _ in results if rc != 0]
    assert not bad, bad
    exe = tmp_path / "cusp_api_tests_nvcc"
    libdir = os.path.join(ROOT, "cusp_autotuned_b200")
    subprocess.check_call(["nvcc", "-o", str(exe)] + [o for *_, o in results] +
                          ["-L", libdir, "-lb200sp", "-Xlinker", "-rpath," + libdir], stderr=subprocess.DEVNULL)
    p = subprocess.run([str(exe), "--host-only"], capture_output=True, text=True, timeout=300)
    passed, failed, skipped = _summary(p.stdout)
    assert p.returncode == 0 and failed == 0 and passed >= 40, p.stdout[-3000:]
