"""Convergence diagnostics across chains: Gelman-Rubin R-hat of the tallied indicators.

The device pools its tallies over the chains of one run, so "a chain" here is a GROUP of chains: G independent runs of
the same problem (seed offsets keep their RNG streams apart), each giving per-indicator counts out of n = window length x
chains-per-group samples.  For an indicator (fragment f takes alignment j / is unmapped; cluster c has support k) the
within-group variance of a 0/1 sample is exactly p_g (1 - p_g), whatever the autocorrelation, so

    W = mean_g p_g (1 - p_g),   B = n * var_g(p_g),   R-hat = sqrt(((n - 1) / n * W + B / n) / W)

is the classic statistic with each group's pooled sample in the role of one chain.  R-hat close to 1 says the groups
agree to within sampling noise; indicators that (almost) never or always occur are skipped (W ~ 0).
"""
from __future__ import annotations

import numpy as np


def rhat_from_group_counts(counts, n_samples: int, min_var: float = 1e-6):
    """counts: array [G, K] of indicator counts per group (G >= 2), n_samples per group.  Returns R-hat per indicator
    (nan where the indicator is constant in every group)."""
    c = np.asarray(counts, np.float64)
    if c.ndim != 2 or c.shape[0] < 2:
        raise ValueError("need counts of at least two groups")
    n = float(n_samples)
    p = c / n
    W = (p * (1 - p)).mean(0)
    B = n * p.var(0, ddof=1)
    out = np.full(c.shape[1], np.nan)
    ok = W > min_var
    out[ok] = np.sqrt(((n - 1) / n * W[ok] + B[ok] / n) / W[ok])
    return out


def run_groups(device_problem, mode, num_iters: int, chains_per_group: int, groups: int = 4, seed: int = 123456, **kw):
    """G independent runs (group g uses seeds seed + g * chains_per_group ...), returning the RunOutputs."""
    return [device_problem.run(mode, num_iters, n_chains=chains_per_group, seed=seed + g * chains_per_group, per_chain=False,
                               want_state=False, **kw) for g in range(groups)]


def rhat(device_problem, mode, num_iters: int, chains_per_group: int, groups: int = 4, burnin: int | None = None,
         seed: int = 123456, **kw):
    """R-hat of every alignment / "unmapped" indicator and of every cluster-support bin.  Returns a dict with the arrays
    `alignments` [n_aln], `unmapped` [n_frag], `supports` [n_hist] and `max` (over the indicators that vary)."""
    from ._abi import reference_burnin
    b = reference_burnin(num_iters) if burnin is None else burnin
    runs = run_groups(device_problem, mode, num_iters, chains_per_group, groups, seed, burnin=b, **kw)
    win_base = max(num_iters - b + 1, 0)
    n = (num_iters - win_base) * chains_per_group
    out = {"alignments": rhat_from_group_counts([r.aln_count for r in runs], n),
           "unmapped": rhat_from_group_counts([r.frag_err_count for r in runs], n),
           "supports": rhat_from_group_counts([r.cluster_hist for r in runs], n)}
    vals = np.concatenate([v[np.isfinite(v)] for v in out.values()])
    out["max"] = float(vals.max()) if vals.size else float("nan")
    out["samples_per_group"] = n
    return out
