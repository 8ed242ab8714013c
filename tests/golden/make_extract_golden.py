"""Generates tests/golden/extract_golden.npz by running the REFERENCE's own landmark functions
(imported from /root/reference through ref_import.py) on seeded inputs. Run in the build
container only:   python tests/golden/make_extract_golden.py
Reference symbols exercised: evaluate_cv.get_mean_rvip_2d (:418-442),
get_ip_from_rvip_mask_3d (:389-416), get_angle2x (:508-536), get_dist (:538-545),
get_distances (:549-561). The threshold -> label-map step (predict_model.py:153-156) sits inside
pred_fold, which needs TF + files; its four lines are applied verbatim below."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference_eval  # noqa: E402


def enc(lst):
    """list of [y,x]/None -> float64 [Z,2] with NaN for None."""
    return np.array([[np.nan, np.nan] if p is None else [float(p[0]), float(p[1])] for p in lst], np.float64)


def make_heat(rng, Z, H, W, kind):
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    heat = np.zeros((Z, H, W, 2), np.float64)
    for z in range(Z):
        for c in range(2):
            present = True
            if kind == 'ragged':
                present = rng.random() > 0.35
            if kind == 'empty':
                present = False
            cy, cx = rng.uniform(4, H - 4), rng.uniform(4, W - 4)
            if kind == 'overlap':
                cy, cx = H / 2 + c * 1.5, W / 2 + c * 1.0
            s = rng.uniform(1.5, 3.5)
            amp = rng.uniform(0.7, 1.0) if present else rng.uniform(0.05, 0.45)
            heat[z, :, :, c] = amp * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
            if kind == 'two_blobs' and c == 0:
                heat[z, :, :, c] += 0.9 * np.exp(-((yy - (H - cy)) ** 2 + (xx - (W - cx)) ** 2) / (2 * 2.0 ** 2))
        heat[z] += 0.02 * rng.random((H, W, 2))
    if kind == 'exact_half':
        heat[:, ::3, ::5, 0] = 0.5          # == thr: strict '>' must exclude these
        heat[:, 1::4, 2::7, 1] = np.nextafter(np.float32(0.5), np.float32(1))
    if kind == 'nan':
        heat[0, 3, 3, 0] = np.nan
        heat[1, 5, 6, 1] = np.nan
    return heat.astype(np.float32)


def main():
    ev = import_reference_eval()
    rng = np.random.default_rng(20211008)
    out = {}
    cases = [('plain', 6, 32, 32), ('ragged', 9, 48, 40), ('empty', 3, 16, 16), ('overlap', 4, 32, 32),
             ('two_blobs', 5, 40, 56), ('exact_half', 3, 24, 24), ('nan', 3, 16, 24), ('plain', 16, 64, 64)]
    names = []
    for i, (kind, Z, H, W) in enumerate(cases):
        preds = make_heat(rng, Z, H, W, kind)
        # ---- predict_model.py:153-156 (verbatim semantics)
        preds_flat = np.zeros((preds.shape[:-1]))
        preds_flat[preds[..., 0] > 0.5] = 1
        preds_flat[preds[..., 1] > 0.5] = 2
        key = 'case%d_%s' % (i, kind)
        names.append(key)
        out[key + '/heat'] = preds
        out[key + '/labels'] = preds_flat.astype(np.uint8)
        for both in (True, False):
            a, b = ev.get_ip_from_rvip_mask_3d(preds_flat.astype(np.uint8), keepdim=True, both_only=both)
            out[key + '/ant_both%d' % both] = enc(a)
            out[key + '/inf_both%d' % both] = enc(b)
        a, b = ev.get_ip_from_rvip_mask_3d(preds_flat.astype(np.uint8), keepdim=False, both_only=True)
        out[key + '/ant_nokeep'] = enc(a) if len(a) else np.zeros((0, 2))
        out[key + '/inf_nokeep'] = enc(b) if len(b) else np.zeros((0, 2))
        # angle / distance on the both-only points (NaN where undefined)
        a, b = ev.get_ip_from_rvip_mask_3d(preds_flat.astype(np.uint8), keepdim=True, both_only=True)
        ang = [np.nan if (p is None or q is None) else ev.get_angle2x(p, q) for p, q in zip(a, b)]
        dst = [np.nan if (p is None or q is None) else ev.get_dist(p, q) for p, q in zip(a, b)]
        out[key + '/angle'] = np.array(ang, np.float64)
        out[key + '/dist'] = np.array(dst, np.float64)
    # SURVEY 8c hand-checkable example
    m = np.zeros((32, 32), np.uint8)
    m[5:8, 10:13] = 1
    m[20:22, 4:9] = 2
    a, b = ev.get_mean_rvip_2d(m)
    out['hand/mask'] = m
    out['hand/ant'] = np.array(a)
    out['hand/inf'] = np.array(b)
    out['hand/angle'] = np.array(ev.get_angle2x(a, b))
    out['hand/dist'] = np.array(ev.get_dist(a, b))
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'extract_golden.npz'), **out)
    print('wrote', len(out), 'arrays')


if __name__ == '__main__':
    main()
