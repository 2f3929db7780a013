"""Golden vectors transcribed from the reference's own unit tests
(/root/reference/testing/*.cu).  Data only; each block cites its source."""
import numpy as np

X = -1  # cusp::ell_matrix::invalid_index

# ---- testing/multiply.cu:441-511 (TestSparseMatrixVectorMultiply) and
#      testing/ktt.cu:214-282: dense inputs A..F; G,H = poisson5pt(4,6), (8,3).
#      x[i] = i % 10, y pre-filled with 10, expected = dense product (exact).
MULTIPLY_DENSE = {
    "A": np.array([[13, 80, 0, 0], [0, 27, 0, 0], [55, 0, 24, 42], [0, 69, 0, 83], [0, 0, 27, 0]], float),
    "B": np.array([[0, 2, 3, 4], [5, 0, 0, 8]], float),
    "C": np.array([[0, 0], [3, 5]], float),
    "D": np.array([[2], [3]], float),
    "E": np.array([[0, 0], [0, 0]], float),
    "F": np.array([[0, 1.5, 3.0], [0.5, 0, 0]], float),
}
MULTIPLY_POISSON = {"G": (4, 6), "H": (8, 3)}

# ---- testing/convert.cu:63-200: the 4x4 / 7-entry conversion example in every format
CONVERT_CSR = dict(format="csr", num_rows=4, num_cols=4, num_entries=7,
                   row_offsets=np.array([0, 2, 3, 6, 7], np.int32),
                   column_indices=np.array([0, 1, 2, 0, 2, 3, 1], np.int32),
                   values=np.array([10.25, 11.00, 12.50, 13.75, 14.00, 15.25, 16.50], np.float32))
CONVERT_COO = dict(format="coo", num_rows=4, num_cols=4, num_entries=7,
                   row_indices=np.array([0, 0, 1, 2, 2, 2, 3], np.int32),
                   column_indices=np.array([0, 1, 2, 0, 2, 3, 1], np.int32),
                   values=np.array([10.25, 11.00, 12.50, 13.75, 14.00, 15.25, 16.50], np.float32))
# dia.resize(4,4,7,3,1): alignment 1 -> pitch 4  (convert.cu:118-140, :405-441)
CONVERT_DIA = dict(format="dia", num_rows=4, num_cols=4, num_entries=7, pitch=4,
                   diagonal_offsets=np.array([-2, 0, 1], np.int32),
                   values=np.array([0, 0, 13.75, 16.50, 10.25, 0, 14.00, 0, 11.00, 12.50, 15.25, 0], np.float32))
# ell.resize(4,4,7,3,1)  (convert.cu:142-175, :443-497)
CONVERT_ELL = dict(format="ell", num_rows=4, num_cols=4, num_entries=7, num_cols_per_row=3, pitch=4,
                   column_indices=np.array([0, 2, 0, 1, 1, X, 2, X, X, X, 3, X], np.int32),
                   values=np.array([10.25, 12.50, 13.75, 16.50, 11.00, 0, 14.00, 0, 0, 0, 15.25, 0], np.float32))
# hyb.resize(4,4,4,3,1,1): 1 ELL column + 3 COO entries (convert.cu:177-200)
CONVERT_HYB = dict(format="hyb", num_rows=4, num_cols=4, num_entries=7,
                   ell=dict(format="ell", num_rows=4, num_cols=4, num_entries=4, num_cols_per_row=1, pitch=4,
                            column_indices=np.array([0, 2, 0, 1], np.int32),
                            values=np.array([10.25, 12.50, 13.75, 16.50], np.float32)),
                   coo=dict(format="coo", num_rows=4, num_cols=4, num_entries=3,
                            row_indices=np.array([0, 2, 2], np.int32),
                            column_indices=np.array([1, 2, 3], np.int32),
                            values=np.array([11.00, 14.00, 15.25], np.float32)))
CONVERT_DENSE = np.array([[10.25, 11.00, 0, 0], [0, 0, 12.50, 0], [13.75, 0, 14.00, 15.25], [0, 16.50, 0, 0]],
                         np.float32)

# ---- testing/poisson.cu:6-93: dense images of the gallery operators
POISSON5_2x3 = np.array([[4, -1, -1, 0, 0, 0], [-1, 4, 0, -1, 0, 0], [-1, 0, 4, -1, -1, 0],
                         [0, -1, -1, 4, 0, -1], [0, 0, -1, 0, 4, -1], [0, 0, 0, -1, -1, 4]], float)
POISSON9_2x3 = np.array([[8, -1, -1, -1, 0, 0], [-1, 8, -1, -1, 0, 0], [-1, -1, 8, -1, -1, -1],
                         [-1, -1, -1, 8, -1, -1], [0, 0, -1, -1, 8, -1], [0, 0, -1, -1, -1, 8]], float)
POISSON7_2x2x2 = np.array([[6, -1, -1, 0, -1, 0, 0, 0], [-1, 6, 0, -1, 0, -1, 0, 0],
                           [-1, 0, 6, -1, 0, 0, -1, 0], [0, -1, -1, 6, 0, 0, 0, -1],
                           [-1, 0, 0, 0, 6, -1, -1, 0], [0, -1, 0, 0, -1, 6, 0, -1],
                           [0, 0, -1, 0, -1, 0, 6, -1], [0, 0, 0, -1, 0, -1, -1, 6]], float)
POISSON27_2x2x2 = np.full((8, 8), -1.0) + 27.0 * np.eye(8)

# ---- testing/format_utils.cu:13-75
OFFSETS = np.array([0, 0, 0, 1, 1, 2, 5, 10], np.int32)
INDICES = np.array([2, 4, 5, 5, 5, 6, 6, 6, 6, 6], np.int32)

# ---- testing/blas.cu:97-142 (axpby), :287-352 (dot/dotc), :434-453 (nrm2), axpy :60-92
BLAS_AXPBY = dict(x=[7.0, 5.0, 4.0, -3.0], y=[0.0, -2.0, 0.0, 5.0], alpha=2.0, beta=1.0, z=[14.0, 8.0, 8.0, -1.0])
BLAS_AXPY = dict(x=[7.0, 5.0, 4.0, -3.0], y=[0.0, -2.0, 0.0, 5.0], alpha=2.0, out=[14.0, 8.0, 8.0, -1.0])
BLAS_DOT = dict(x=[7.0, 5.0, 4.0, -3.0, 0.0, 4.0], y=[0.0, -2.0, 0.0, 5.0, 6.0, 1.0], result=-21.0)
BLAS_NRM2 = dict(x=[7.0, 5.0, 4.0, -3.0, 0.0, 1.0], result=10.0)

# ---- testing/cg.cu:46-99: poisson5pt(10,10) fp32, b = 1, x0 = 0, monitor(b, 20, 1e-4):
#      must reach ||b - A x|| < 1e-4 ||b||;  zero-residual case: diag(8,4), x = 1 ->
#      iteration_count() == 0 and converged.
CG_CASE = dict(grid=(10, 10), limit=20, rel=1e-4)

# ---- testing/ktt.cu:274-281: all-ones banded matrices, every configuration
KTT_BANDED = [(4096, 4096, 1, 1024), (4096, 2048, 1, 1024), (2048, 4096, 1, 1024)]
