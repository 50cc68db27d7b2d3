"""Round-2 study (CPU, NumPy): does a STATIC per-row exponent e_i = ceil(log2 sqrt(A_ii)) (valid for every entry of row
i of L because sum_j L_ij^2 = A_ii) keep the digit-sliced trailing updates as accurate as the per-update row-max
exponent that slice_rows_kernel computes today?  If so the panel kernel can emit the int8 digit planes itself and the
slicing pre-pass (0.5 ms of the 8.66 ms C2 evaluation) disappears.  Emulates the recursive factorisation of potrf.cu
with the digit-sliced update of gemm_tc.cu for K >= KMIN."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import gp_oracle as O

S, NB = 7, 128


def digits(P, e):
    v = P * np.exp2(6.0 - e)[:, None]
    D = np.empty((S,) + P.shape, dtype=np.int64)
    for s in range(S):
        d = np.rint(v)
        d = np.clip(d, -64, 64)
        v = (v - d) * 128.0
        D[s] = d
    return D


def sliced_update(C, P, n, e_rows):
    """C[m, n] -= P P[0:n]^T with digit slicing; e_rows: exponents per row of P (None = per-update row max)."""
    m = P.shape[0]
    if e_rows is None:
        mx = np.abs(P).max(axis=1)
        e_rows = np.where(mx > 0, np.floor(np.log2(np.where(mx > 0, mx, 1.0))) + 1, 0)
    D = digits(P, e_rows)
    rs = np.exp2(e_rows - 6.0)
    acc = np.zeros((m, n))
    for g in range(S):
        a = np.zeros((m, n), dtype=np.int64)
        for s in range(g + 1):
            a += D[s] @ D[g - s][:n].T
        acc += a * 2.0 ** (-7 * g)
    C -= acc * rs[:, None] * rs[None, :n]


def potrf_rec(A, n, e_static, kmin, off=0):
    rows = A.shape[0]
    if n <= NB:
        A[:n, :n] = np.linalg.cholesky(A[:n, :n])
        if rows > n:
            A[n:, :n] = np.linalg.solve(A[:n, :n], A[n:, :n].T).T
        return
    n1 = ((n // NB + 1) // 2) * NB
    potrf_rec(A, n1, e_static, kmin, off)
    P = A[n1:, :n1]
    C = A[n1:, n1:n]
    if n1 >= kmin:
        sliced_update(C, P, n - n1, None if e_static is None else e_static[off + n1:off + rows])
    else:
        C -= P @ P[:n - n1].T
    potrf_rec(A[n1:, n1:], n - n1, e_static, kmin, off + n1)


def run(N, kmin, static):
    d = O.make_data(2, N, 8, 1)
    K = O.Matern52(lengthscales=np.sqrt(8.0))(d["X"]) + 0.1 * np.eye(N)
    Lx = np.linalg.cholesky(K)
    A = K.copy()
    e = np.ceil(np.log2(np.sqrt(np.diag(K)))) + 0.0 if static else None
    if static:
        e = e + 1   # digits must satisfy |x 2^(6-e)| <= 64: one extra bit of headroom over sqrt(A_ii)
    potrf_rec(A, N, e, kmin)
    L = np.tril(A)
    return np.abs(L - Lx).max() / np.abs(Lx).max(), abs(np.sum(np.log(np.diag(L))) - np.sum(np.log(np.diag(Lx))))


if __name__ == "__main__":
    for N in (1024, 2048):
        for static in (False, True):
            err, dlogdet = run(N, 256, static)
            print(f"N={N} kmin=256 scale={'static sqrt(A_ii)' if static else 'per-update row max'}: max|dL|/max|L| = {err:.2e}, |d sum log diag| = {dlogdet:.2e}")
