#!/usr/bin/env python
"""Group-conditional coverage with BatchMVP and multicalibration of a biased probability estimate (the reference's
``examples/multivalid_coverage.pct.py`` setting): conformal thresholds that cover 90 % of the scores overall can badly
miss sub-populations; BatchMVP patches the thresholds until every (group, threshold bucket) cell is covered at the
target rate.  The per-round statistics of all cells come from one kernel pass (fb200_multivalid_stats)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortuna_b200.conformal import BatchMVPConformalRegressor, Multicalibrator  # noqa: E402


def main():
    r = np.random.default_rng(0)
    n = 20000
    # two overlapping sub-populations with different score scales, and the group of everybody (a point that belongs to no
    # group is never patched)
    groups = np.stack([r.random(n) < 0.5, r.random(n) < 0.2, np.ones(n, bool)], axis=1)
    noise = 0.05 + 0.15 * groups[:, 0] + 0.25 * groups[:, 1]
    scores = np.clip(np.abs(r.standard_normal(n)) * noise, 0, 1).astype(np.float32)
    test = slice(15000, None)
    coverage = 0.9
    q = np.quantile(scores[:15000], coverage)
    print(f"one marginal threshold {q:.3f}: coverage overall {np.mean(scores[test] <= q):.3f}, "
          + ", ".join(f"group {g} {np.mean(scores[test][groups[test, g]] <= q):.3f}" for g in range(2)))
    bm = BatchMVPConformalRegressor(seed=0)
    th, status = bm.calibrate(scores=scores[:15000], groups=groups[:15000], test_groups=groups[test], coverage=coverage,
                              n_buckets=50, n_rounds=200, eta=0.5)
    th = th.cpu().numpy()
    print(f"BatchMVP ({status['n_rounds']} patches, converged={status['converged']}): coverage overall "
          f"{np.mean(scores[test] <= th):.3f}, "
          + ", ".join(f"group {g} {np.mean(scores[test][groups[test, g]] <= th[groups[test, g]]):.3f}" for g in range(2)))

    # multicalibration: a probability estimate that is biased inside one group
    p_true = r.random(n).astype(np.float32)
    y = (r.random(n) < p_true).astype(np.float32)
    p_hat = np.clip(p_true + 0.2 * groups[:, 1], 0, 1).astype(np.float32)  # over-confident inside group 1
    mc = Multicalibrator(seed=0)
    before = mc.calibration_error(y[test], groups[test], p_hat[test]).cpu().numpy()
    fixed, st = mc.calibrate(scores=y[:15000], groups=groups[:15000], values=p_hat[:15000], test_groups=groups[test],
                             test_values=p_hat[test], n_buckets=20, n_rounds=100, eta=0.5)
    after = mc.calibration_error(y[test], groups[test], fixed).cpu().numpy()
    print(f"multicalibrator ({st['n_rounds']} patches): per-group calibration error {np.round(before, 4)} -> {np.round(after, 4)}")


if __name__ == "__main__":
    main()
