"""CPU: the landmark-metrics oracle against golden vectors produced by the reference's own functions."""
import os

import numpy as np

from oracle import metrics_ref as M

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'metrics_golden.npz')


def _cases():
    g = np.load(GOLD, allow_pickle=False)
    for c in g['cases']:
        yield str(c), {k[len(str(c)) + 1:]: g[k] for k in g.files if k.startswith(str(c) + '_')}


def _eq(a, b):
    return np.allclose(a, b, rtol=1e-12, atol=1e-12, equal_nan=True)


def test_metrics_oracle_matches_reference_golden():
    n = 0
    for name, c in _cases():
        spacing, thr, dim = c['params']
        assert _eq(M.angles(c['gt_ant'], c['gt_inf']), c['angle_gt']), name
        assert _eq(M.angles(c['pr_ant'], c['pr_inf']), c['angle_pr']), name
        for lm in ('ant', 'inf'):
            gt, pr = c['gt_' + lm], c['pr_' + lm]
            assert _eq(M.distances(gt, pr, spacing), c['dist_' + lm]), name
            assert _eq(M.distances(gt, pr, spacing, thr), c['dist_thr_' + lm]), name
            assert _eq(M.distances_upper_bound(gt, pr, spacing, dim), c['ub_' + lm]), name
        for i, lm in enumerate(('ant', 'inf')):
            tpr, ppv, *_ = M.tpr_ppv(c['gt_' + lm], c['pr_' + lm], thr, spacing)
            assert abs(tpr - c['tpr'][i]) < 1e-15 and abs(ppv - c['ppv'][i]) < 1e-15, name
        for tag in ('gt', 'pr'):
            ma, mi = M.mean_ip(c[tag + '_ant'], c[tag + '_inf'])
            assert _eq(ma, c['mean_%s_ant' % tag]) and _eq(mi, c['mean_%s_inf' % tag]), name
        n += 1
    assert n == 5
