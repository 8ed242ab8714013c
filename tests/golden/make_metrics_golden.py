"""Generates tests/golden/metrics_golden.npz by running the REFERENCE's own landmark-metric functions (imported
from /root/reference through ref_import.py) on seeded point lists.  Run in the build container only:
    python tests/golden/make_metrics_golden.py
Reference symbols exercised (src/models/evaluate_cv.py): get_angle2x :508-536, get_distances :549-561,
get_distances_upper_bound :572-595, calc_tpr_thresh :267-308, calc_ppv_thresh :311-353.  calc_mean_ip :113-120 uses
np.NaN (removed in numpy 2): its three lines are applied verbatim below with np.nan."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference_eval  # noqa: E402


def to_list(a):
    return [None if not np.isfinite(p).all() else [float(p[0]), float(p[1])] for p in a]


def enc(v):
    return np.array([np.nan if x is None else float(x) for x in v], np.float64)


def main():
    ev = import_reference_eval()
    rng = np.random.default_rng(20211009)
    out, names = {}, []
    for i, (Z, p_missing, dim) in enumerate([(12, 0.0, 224), (16, 0.3, 224), (9, 0.6, 256), (5, 1.0, 224), (24, 0.2, 128)]):
        pts = {}
        for k in ('gt_ant', 'gt_inf', 'pr_ant', 'pr_inf'):
            a = rng.uniform(5, dim - 5, size=(Z, 2))
            if k.startswith('pr'):
                a = pts['gt' + k[2:]] + rng.normal(0, 6, size=(Z, 2))     # predictions near the ground truth
            a[rng.random(Z) < p_missing] = np.nan
            pts[k] = a
        gt, pr = (to_list(pts['gt_ant']), to_list(pts['gt_inf'])), (to_list(pts['pr_ant']), to_list(pts['pr_inf']))
        spacing, thr = float(rng.uniform(0.8, 1.6)), float(rng.uniform(4, 12))
        o = {}
        # points may be None in the lists; get_angle2x expects finite tuples or NaN-coded ones
        o['angle_gt'] = enc([ev.get_angle2x(a, b) if a is not None and b is not None else None for a, b in zip(*gt)])
        o['angle_pr'] = enc([ev.get_angle2x(a, b) if a is not None and b is not None else None for a, b in zip(*pr)])
        da, di = ev.get_distances(gt, pr, spacing=spacing)
        o['dist_ant'], o['dist_inf'] = enc(da), enc(di)
        da, di = ev.get_distances(gt, pr, spacing=spacing, threshold=thr)
        o['dist_thr_ant'], o['dist_thr_inf'] = enc(da), enc(di)
        da, di = ev.get_distances_upper_bound(gt, pr, spacing=spacing, dim=dim)
        o['ub_ant'], o['ub_inf'] = enc(da), enc(di)
        o['tpr'] = np.array(ev.calc_tpr_thresh(gt, pr, thresh=thr, spacing=spacing), np.float64)
        o['ppv'] = np.array(ev.calc_ppv_thresh(gt, pr, thresh=thr, spacing=spacing), np.float64)
        # calc_mean_ip (evaluate_cv.py:113-120), np.NaN -> np.nan
        for tag, lst in (('gt', gt), ('pr', pr)):
            mant, minf = np.nan, np.nan
            ants, infs = [e for e in lst[0] if e is not None], [e for e in lst[1] if e is not None]
            if len(ants) > 0 and len(infs) > 0:
                mant, minf = np.array(ants).mean(axis=0), np.array(infs).mean(axis=0)
            o['mean_%s_ant' % tag] = np.broadcast_to(np.asarray(mant, np.float64), (2,)).copy()
            o['mean_%s_inf' % tag] = np.broadcast_to(np.asarray(minf, np.float64), (2,)).copy()
        for k, v in pts.items():
            out['c%d_%s' % (i, k)] = v
        for k, v in o.items():
            out['c%d_%s' % (i, k)] = v
        out['c%d_params' % i] = np.array([spacing, thr, dim], np.float64)
        names.append('c%d' % i)
    out['cases'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'metrics_golden.npz'), **out)
    print('wrote', len(out), 'arrays')


if __name__ == '__main__':
    main()
