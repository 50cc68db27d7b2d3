"""NumPy model of the fp64 -> int8 digit slicing of gpflow_b200/csrc/gemm_tc.cu (slice_rows_kernel + the weighted
recombination of the tcgen05 int32 accumulators): checks, without a GPU, the error bound DESIGN.md 4.3 states for the
tensor-core trailing update, the exactness of the digit expansion, and the int32 headroom of the accumulators."""
import numpy as np
import pytest


def slice_rows(P: np.ndarray, S: int):
    """Per row: e = ilogb(max|row|) + 1, digits d_s in [-64, 64] of v = x 2^(6-e) with remainder rescaled by 128
    (slice_rows_kernel, gemm_tc.cu:47-85).  Returns digits [S, m, K] (int64) and rowscale [m] = 2^(e-6)."""
    mx = np.abs(P).max(axis=1)
    e = np.where(mx > 0, np.floor(np.log2(np.where(mx > 0, mx, 1.0))).astype(np.int64) + 1, 0)
    v = P * np.exp2(6.0 - e)[:, None]
    digits = np.empty((S,) + P.shape, dtype=np.int64)
    for s in range(S):
        d = np.rint(v)
        assert np.abs(d).max() <= 64
        v = (v - d) * 128.0          # exact in fp64
        digits[s] = d.astype(np.int64)
    return digits, np.exp2(e - 6.0), e


def syrk_model(P: np.ndarray, S: int) -> np.ndarray:
    """sum_{s+t<S} 2^(-7(s+t)) D_s D_t^T, recombined with the row scales (syrk_i8_kernel epilogue)."""
    D, rs, _ = slice_rows(P, S)
    m = P.shape[0]
    out = np.zeros((m, m))
    for g in range(S):
        acc = np.zeros((m, m), dtype=np.int64)
        for s in range(g + 1):
            acc += D[s] @ D[g - s].T                      # exact integer accumulation (int32 on the tensor cores)
        assert np.abs(acc).max() < 2 ** 31, "int32 accumulator would overflow"
        out += acc.astype(np.float64) * 2.0 ** (-7 * g)
    return out * rs[:, None] * rs[None, :]


@pytest.mark.parametrize("S", [6, 7, 8])
@pytest.mark.parametrize("K", [512, 4096])
def test_digit_sliced_syrk_error_bound(S, K):
    rng = np.random.default_rng(S * 1000 + K)
    m = 96
    P = rng.standard_normal((m, K)) * np.exp2(rng.integers(-20, 20, size=(m, 1)))   # rows of very different scale
    got = syrk_model(P, S)
    Pl = P.astype(np.longdouble)
    exact = (Pl @ Pl.T).astype(np.float64)
    _, _, e = slice_rows(P, S)
    bound = 1.5 * K * S * 2.0 ** (-7 * S) * np.exp2(e)[:, None] * np.exp2(e)[None, :] + 4 * np.finfo(float).eps * np.abs(exact)
    assert np.all(np.abs(got - exact) <= bound)
    if S == 7:   # the default: ~1e-11 relative to the row scales at K = 4096
        rel = np.abs(got - exact) / (np.exp2(e)[:, None] * np.exp2(e)[None, :])
        assert rel.max() < 4096 * 7 * 2.0 ** -49 * 1.5


def test_digit_expansion_is_exact_up_to_the_last_digit():
    rng = np.random.default_rng(0)
    P = rng.standard_normal((8, 64))
    for S in (6, 7, 8):
        D, rs, _ = slice_rows(P, S)
        recon = sum(D[s] * 2.0 ** (-7 * s) for s in range(S)) * rs[:, None]
        assert np.abs(recon - P).max() <= 0.5 * 2.0 ** (-7 * (S - 1)) * rs.max() * 1.0000001


def test_int32_headroom_at_the_largest_k():
    """Worst case |digit| = 64 everywhere and (g+1) <= S digit pairs per accumulator: 64*64*K*S must stay below 2^31,
    i.e. K <= 2^19 / S (74898 at S = 7).  potrf.cu::trailing_update falls back to the DMMA kernel beyond that."""
    assert 64 * 64 * 4096 * 7 < 2 ** 31                  # BASELINE config 2: K <= N/2 = 4096
    assert 64 * 64 * 65536 * 8 >= 2 ** 31 > 64 * 64 * 65535 * 8
