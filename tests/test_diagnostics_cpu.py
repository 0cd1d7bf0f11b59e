"""R-hat across chain groups (multibreak_sv_b200.diagnostics): the statistic itself on synthetic Bernoulli samples."""
import numpy as np
import pytest


def test_rhat_is_one_for_agreeing_groups_and_large_for_disagreeing_ones():
    from multibreak_sv_b200 import diagnostics
    rng = np.random.default_rng(1)
    n = 20000
    p = np.array([0.5, 0.1, 0.9, 0.02, 0.0, 1.0])
    same = rng.binomial(n, p, size=(6, p.size))
    r = diagnostics.rhat_from_group_counts(same, n)
    assert np.isnan(r[4]) and np.isnan(r[5])                          # constant indicators are skipped
    assert np.all(np.abs(r[:4] - 1) < 0.002)
    # one group stuck elsewhere: B/n dominates
    stuck = same.copy()
    stuck[0, 0] = rng.binomial(n, 0.8)                                # var_g(p) = 0.015, W = 0.235: sqrt(1 + 0.064)
    r2 = diagnostics.rhat_from_group_counts(stuck, n)
    assert 1.025 < r2[0] < 1.04 and np.all(np.abs(r2[1:4] - 1) < 0.002)
    two = diagnostics.rhat_from_group_counts([[rng.binomial(n, 0.2)], [rng.binomial(n, 0.8)]], n)
    assert two[0] > 1.4                                               # two groups in different modes
    # closed form on a hand-made case: p = (0.4, 0.6), n = 10 -> W = 0.24, B = 10 * 0.02 = 0.2
    r3 = diagnostics.rhat_from_group_counts([[4], [6]], 10)
    assert abs(r3[0] - np.sqrt((0.9 * 0.24 + 0.02) / 0.24)) < 1e-12
    with pytest.raises(ValueError):
        diagnostics.rhat_from_group_counts([[1, 2]], 10)
