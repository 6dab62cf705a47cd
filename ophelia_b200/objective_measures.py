"""Validation scores of the reference's `objective_measures.py` (used by `train.py:37-46` after every epoch).

The reference delegates to the third-party package `mcd` (github.com/MattShannon/mcd, un-pinned in requirements.txt and
absent from this image): `mcd.metrics_fast.logSpecDbDist` and `mcd.dtw.dtw`.  Their published definitions are restated
here with numpy:

* logSpecDbDist(x, y) = (10 / ln 10) * sqrt(2) * ||x - y||_2                         (log-spectral distance in dB)
* dtw(xs, ys, cost): minimum over monotone alignment paths with steps (1,0), (0,1), (1,1) of the summed local costs.
"""
import math

import numpy as np

LOG_SPEC_DB_CONST = 10.0 / math.log(10.0) * math.sqrt(2.0)


def logSpecDbDist(x, y):
    diff = np.asarray(x, np.float64) - np.asarray(y, np.float64)
    return LOG_SPEC_DB_CONST * math.sqrt(float(np.inner(diff, diff)))


def _cost_matrix(nat, synth):
    """All pairwise logSpecDbDist values [len(nat), len(synth)] (one GEMM instead of len*len python calls)."""
    nn = (nat * nat).sum(1)[:, None]
    ss = (synth * synth).sum(1)[None, :]
    d2 = np.maximum(nn + ss - 2.0 * nat @ synth.T, 0.0)
    return LOG_SPEC_DB_CONST * np.sqrt(d2)


def dtw_min_cost(cost):
    """Cumulative cost of the best path through `cost` (steps down, right, diagonal).  Each row's recurrence
    cur[j] = c[j] + min(prev[j-1], prev[j], cur[j-1]) is a min-plus scan, evaluated with two numpy scans."""
    cur = np.cumsum(cost[0])
    for i in range(1, cost.shape[0]):
        c = cost[i]
        best_prev = cur.copy()
        best_prev[1:] = np.minimum(cur[1:], cur[:-1])
        run = np.cumsum(c)
        cur = run + np.minimum.accumulate(c + best_prev - run)
    return float(cur[-1])


def compute_dtw_error(reference, predictions):
    """objective_measures.py:12-24: total DTW cost over total natural frames."""
    cost_tot, frames_tot = 0.0, 0
    for nat, synth in zip(reference, predictions):
        nat, synth = np.asarray(nat, np.float64), np.asarray(synth, np.float64)
        cost_tot += dtw_min_cost(_cost_matrix(nat, synth))
        frames_tot += len(nat)
    mean_score = cost_tot / frames_tot
    print('overall LSD = %f (%s frames nat/synth)' % (mean_score, frames_tot))
    return mean_score


def compute_simple_LSD(reference_list, prediction_list):
    """objective_measures.py:26-43: frame-synchronous log-spectral distance (equal lengths required)."""
    cost_tot, frames_tot = 0.0, 0
    for synth, nat in zip(prediction_list, reference_list):
        assert len(synth) == len(nat)
        d = np.asarray(nat, np.float64) - np.asarray(synth, np.float64)
        cost_tot += LOG_SPEC_DB_CONST * float(np.sqrt((d * d).sum(1)).sum())
        frames_tot += len(nat)
    return cost_tot / frames_tot
