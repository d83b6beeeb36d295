"""CPU checks of the two arithmetic facts the sparse ion sub-cycle (csrc/ion.cu) rests on; the kernels themselves are
tested on the GPU against the reference's fingerprints (tests/test_gpu_longrun.py::test_sparse_subcycle_...)."""
import numpy as np


def test_increment_that_rounds_away_at_dt_max_rounds_away_for_every_shorter_substep():
    """k_ion_update marks a chunk CH_UINERT when E + edot*dt_max == E and s + nHdot*dt_max*m_H == s (ionization_update,
    ionrad_3d.c:575-580, evaluated at the longest sub-step still possible in the call).  Rounding is monotone, so the same
    holds for every 0 <= dt <= dt_max: the later updates of that chunk, whatever sub-step the loop picks, change nothing."""
    rng = np.random.default_rng(3)
    n = 400000
    E = 10.0 ** rng.uniform(-12, 6, n)
    # increments around the rounding threshold of E (half an ulp), both signs
    edot = rng.choice([-1.0, 1.0], n) * np.spacing(E) * 10.0 ** rng.uniform(-3, 1, n)
    dt_max = 10.0 ** rng.uniform(-2, 2, n)
    edot = edot / dt_max
    m_H = 1.67e-24
    stays = (E + edot * dt_max) == E
    assert 0.2 < stays.mean() < 0.8, "the sample must straddle the threshold"
    for frac in (1.0, 0.999999, 0.75, 0.5, 0.1, 1e-3, 1e-9, 0.0):
        dt = dt_max * frac
        assert np.all(dt <= dt_max)
        assert np.all((E + edot * dt)[stays] == E[stays]), frac
    s = E * m_H
    nHdot = edot / m_H
    stays_s = (s + nHdot * dt_max * m_H) == s
    for frac in (0.999999, 0.5, 1e-3):
        assert np.all((s + nHdot * (dt_max * frac) * m_H)[stays_s] == s[stays_s]), frac


def test_bit_pattern_order_of_nonnegative_doubles():
    """The chunk records hold min dt limits and max wave speeds reduced with redux.sync on the (hi, lo) words of the bit
    patterns (warp_min_ordered / warp_max_ordered); that is the numerical min / max for non-negative, non-NaN doubles -- the
    only values ordered_bits() lets into a record (+0, denormals, DBL_MAX and +inf included)."""
    rng = np.random.default_rng(4)
    x = np.concatenate([10.0 ** rng.uniform(-320, 308, 100000), [0.0, 5e-324, 2.2250738585072014e-308, np.finfo(float).max, np.inf]])
    bits = x.view(np.uint64)
    order = np.argsort(x, kind="stable")
    assert np.all(np.diff(bits[order].astype(np.float64)) >= 0) and np.all(np.diff(x[order]) >= 0)
    for _ in range(200):
        w = rng.choice(len(x), 32)
        hi, lo = (bits[w] >> np.uint64(32)).astype(np.uint32), (bits[w] & np.uint64(0xffffffff)).astype(np.uint32)
        mh = hi.min(); ml = np.where(hi == mh, lo, np.uint32(0xffffffff)).min()
        assert np.array([(np.uint64(mh) << np.uint64(32)) | np.uint64(ml)], dtype=np.uint64).view(np.float64)[0] == x[w].min()
        mh = hi.max(); ml = np.where(hi == mh, lo, np.uint32(0)).max()
        assert np.array([(np.uint64(mh) << np.uint64(32)) | np.uint64(ml)], dtype=np.uint64).view(np.float64)[0] == x[w].max()
    # what ordered_bits() rejects: negatives (incl. -0.0) and NaN have the sign bit set or exceed +inf's pattern
    bad = np.array([-0.0, -1.0, np.nan, -np.inf]).view(np.uint64)
    assert np.all(bad > np.array([np.inf]).view(np.uint64)[0])
