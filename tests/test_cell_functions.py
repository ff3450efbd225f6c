"""CPU mirrors of the two lookup-cell functions of the parallel mode's streaming kernel
(crick_b200/csrc/tdigest_parallel.cu: skey32 / cell_key, cell_lin).  The kernel only needs them to be
monotone non-decreasing in x (edges and values go through the same function, so a value never lands
in a cell left of an edge below it), with -0.0 and +0.0 in the same cell (the reference's `<` on
means treats them as equal, tdigest_stubs.c:116-175)."""
import numpy as np

K_NC = 8192


def skey32(x):
    hi = (np.asarray(x, dtype=np.float64).view(np.int64) >> 32).astype(np.int64)      # signed high word
    m = hi >> 31                                                                        # 0 or -1
    return ((hi ^ (m & 0x7fffffff)) - m).astype(np.int64)


def cell_key(x, kb1, shift):
    return np.clip(((skey32(x) >> 1) - kb1) >> shift, 0, K_NC - 1)


def cell_lin(x, c0, inv):
    t = np.asarray(x, dtype=np.float64) * inv + c0        # (the kernel fuses this: one rounding; both are monotone)
    with np.errstate(invalid="ignore"):
        c = np.floor(t)
    c = np.where(np.isnan(c) | (c < 0), 0, c)             # cvt.rmi.u32.f64 saturates: negative and NaN -> 0
    return np.minimum(c, K_NC - 1).astype(np.int64)


def _sorted_doubles(rng, n):
    x = np.concatenate([rng.normal(size=n), rng.lognormal(0, 20, n), -rng.lognormal(0, 20, n),
                        [0.0, -0.0, 5e-324, -5e-324, 2.2e-308, -2.2e-308, np.inf, -np.inf, 1.0, np.nextafter(1.0, 2.0)]])
    return np.sort(x, kind="stable")


def test_skey32_is_monotone_and_ties_the_zeros():
    rng = np.random.default_rng(0)
    x = _sorted_doubles(rng, 20000)
    k = skey32(x)
    assert (np.diff(k) >= 0).all()
    assert skey32(np.array([0.0]))[0] == skey32(np.array([-0.0]))[0] == 0
    assert skey32(np.array([-5e-324]))[0] <= 0 <= skey32(np.array([5e-324]))[0]
    assert -2**31 < k.min() and k.max() < 2**31           # fits the 32-bit register the kernel uses


def test_cell_key_is_monotone_for_any_base_and_shift():
    rng = np.random.default_rng(1)
    x = _sorted_doubles(rng, 20000)
    for shift in (0, 3, 11, 19):
        for kb1 in (-(2**30), -12345, 0, 987654, 2**29):
            c = cell_key(x, kb1, shift)
            assert (np.diff(c) >= 0).all() and c.min() >= 0 and c.max() <= K_NC - 1
            assert cell_key(np.array([0.0]), kb1, shift)[0] == cell_key(np.array([-0.0]), kb1, shift)[0]


def test_cell_lin_is_monotone_and_saturates():
    rng = np.random.default_rng(2)
    x = _sorted_doubles(rng, 20000)
    for lo, hi in ((-3.0, 3.0), (1e-3, 1e5), (-1e300, 1e300), (5.0, 5.0 + 1e-9)):
        inv = (K_NC - 2) / (hi - lo) if hi > lo else 0.0
        c0 = 1.0 - lo * inv
        c = cell_lin(x, c0, inv)
        assert (np.diff(c) >= 0).all() and c.min() >= 0 and c.max() <= K_NC - 1
    assert cell_lin(np.array([np.nan]), 1.0, 1.0)[0] == 0
