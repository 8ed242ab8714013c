"""Per-kernel-class DRAM traffic of ONE training step from an ncu launch list
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active... --csv of
`python profiles/step_for_ncu.py 1 1`: two steps, the second one is summarised).  Writes the JSON bench.py reads for
`roofline.traffic` (class -> dram_mbytes_per_step, launches, ncu time share, tensor-pipe activity).
usage: python profiles/step_traffic.py <launches.csv> <out.json> [source note]"""
import csv
import json
import re
import sys


def classify(name, after_head):
    if 'c1_fwd' in name or 'wgrad3x3_c1' in name or 'simt' in name:
        return 'conv_cuda_core'
    if 'wgrad3x3' in name:
        return 'conv_wgrad_tcgen05'
    if re.search(r'conv3x3_(row|halo|tc)_kernel', name):
        return 'conv_dgrad_tcgen05' if after_head else 'conv_fwd_tcgen05'
    if 'bn_apply' in name or 'bn_eval' in name:
        return 'bn_forward'
    if 'bn_bwd' in name or 'relu_bwd' in name:
        return 'bn_backward'
    if 'head' in name:
        return 'head_loss'
    if 'adam' in name or 'sgd' in name or 'pack_' in name:
        return 'adam_pack'
    if 'extract' in name or 'label_map' in name:
        return 'extract'
    return 'memset_misc'


def main():
    raw, out = sys.argv[1:3]
    note = sys.argv[3] if len(sys.argv) > 3 else raw
    rows = [r for r in csv.reader(open(raw)) if r and not r[0].startswith('==')]
    hdr = rows[0]
    idx = {n: i for i, n in enumerate(hdr)}
    # long format (one row per metric) or wide format (one column per metric)
    launches = []
    if 'Metric Name' in idx:
        cur = {}
        for r in rows[1:]:
            if len(r) <= idx['Metric Value']:
                continue
            key = r[idx['ID']]
            if not cur or cur['id'] != key:
                cur = {'id': key, 'name': r[idx['Kernel Name']]}
                launches.append(cur)
            v = r[idx['Metric Value']].replace(',', '')
            try:
                v = float(v)
            except ValueError:
                continue
            unit = r[idx['Metric Unit']]
            scale = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0, 'us': 1e3, 'usecond': 1e3, 'ms': 1e6,
                     'msecond': 1e6, 'ns': 1.0, 'nsecond': 1.0, 'second': 1e9}.get(unit, 1.0)
            cur[r[idx['Metric Name']]] = v * scale
    else:
        raise SystemExit('expected the long CSV format of `ncu --csv` (one row per metric)')
    starts = [i for i, l in enumerate(launches) if 'c1_fwd' in l['name']]
    step = launches[starts[-1]:] if starts else launches
    per, after_head = {}, False
    for l in step:
        cls = classify(l['name'], after_head)
        if 'head_kernel' in l['name'] and 'Lb1' in l['name'] or ('head_kernel' in l['name'] and ', 1,' in l['name'][:60]):
            after_head = True
        if 'head_kernel' in l['name']:
            after_head = True
        e = per.setdefault(cls, {'launches': 0, 'ns': 0.0, 'bytes': 0.0, 'pipe': [], 'kernels': {}})
        e['launches'] += 1
        e['ns'] += l.get('gpu__time_duration.sum', 0.0)
        e['bytes'] += l.get('dram__bytes_read.sum', 0.0) + l.get('dram__bytes_write.sum', 0.0)
        p = l.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')
        if p is not None and cls.startswith('conv_') and cls != 'conv_cuda_core':
            e['pipe'].append((p, l.get('gpu__time_duration.sum', 0.0)))
    tot_ns = sum(e['ns'] for e in per.values())
    res = {}
    for cls, e in per.items():
        r = {'dram_mbytes_per_step': round(e['bytes'] / 1e6, 1), 'launches_per_step': e['launches'],
             'ncu_ms_per_step': round(e['ns'] / 1e6, 4), 'ncu_time_share': round(e['ns'] / tot_ns, 4), 'source': note}
        if e['pipe']:
            w = sum(t for _, t in e['pipe'])
            r['tensor_pipe_active_pct_time_weighted'] = round(sum(p * t for p, t in e['pipe']) / max(w, 1e-9), 1)
        res[cls] = r
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
