"""NumPy mirror of the fp64 fast math of the K-build fast path (kbuild.cu::stationary_value4): table-driven exp(-u)
(2^(j/64) table + degree-5 polynomial, two-constant Cody-Waite reduction) and the MUFU-seeded coupled Newton / Heron
square root.  Plain fp64 arithmetic stands in for the device's fused multiply-adds (which only round less), so the
bounds asserted here are upper bounds for the device code."""
import numpy as np

MAGIC = 6755399441055744.0          # 1.5 * 2^52
C64 = 92.33248261689366             # 64 / ln 2
HI, LO = 0.01083042468962958, 6.619564634077006e-12
TAB = np.exp2(np.arange(64) / 64.0)


def fast_exp_neg(u):
    sh = u * (-C64) + MAGIC
    kd = sh - MAGIC
    k = sh.view(np.int64).astype(np.int32).astype(np.int64)      # __double2loint
    rr = kd * (-HI) + (-u)
    rr = kd * (-LO) + rr
    p = rr * 8.3333333333333332e-03 + 4.1666666666666664e-02
    p = p * rr + 1.6666666666666666e-01
    p = p * rr + 0.5
    p = p * rr + 1.0
    p = p * rr + 1.0
    return TAB[k & 63] * p * np.exp2((k >> 6).astype(np.float64))


def fast_sqrt(x, seed_bits=22):
    y0 = 1.0 / np.sqrt(x)
    y0 = y0 * (1.0 + np.random.default_rng(0).uniform(-1, 1, x.shape) * 2.0 ** -seed_bits)   # MUFU.RSQ accuracy
    y0 = y0.astype(np.float32).astype(np.float64)
    g, h = x * y0, 0.5 * y0
    r = 0.5 - g * h
    g = g + g * r
    d = x - g * g
    return g + h * d


def test_constants_are_what_the_comments_say():
    assert abs(C64 - 64 / np.log(2)) < 1e-13
    assert abs((HI + LO) - np.log(2) / 64) < 1e-27 + 2 ** -70
    assert (np.float64(HI).view(np.uint64) & np.uint64((1 << 22) - 1)) == 0      # kd * HI is exact for |kd| < 2^22


def test_table_exp_accuracy():
    rng = np.random.default_rng(1)
    u = np.concatenate([rng.uniform(0, 40, 200000), rng.uniform(0, 700, 200000), [0.0, 1e-300, 1e-9, 707.9]])
    got = fast_exp_neg(u)
    ref = np.exp(-u.astype(np.longdouble)).astype(np.float64)
    rel = np.abs(got - ref) / ref
    assert rel.max() < 4.5e-16, rel.max()          # ~2 ulp worst case without fma; < 1 ulp typical


def test_newton_heron_sqrt_accuracy():
    rng = np.random.default_rng(2)
    x = np.exp(rng.uniform(np.log(5e-36), np.log(1e30), 400000))
    got = fast_sqrt(x)
    ref = np.sqrt(x.astype(np.longdouble)).astype(np.float64)
    rel = np.abs(got - ref) / ref
    assert rel.max() < 3.4e-16, rel.max()          # one Newton step on g, one Heron correction: 1.5 e0^3 + roundings
