"""Condenses `ncu -i <report> --page raw --csv` (one row per captured launch) into the per-launch summary table and the
per-kernel-family traffic file that bench.py reads for `roofline.traffic`.
usage: python profiles/ncu_summary.py <raw.csv> <summary.csv> <traffic.json> [source note]"""
import csv
import json
import re
import sys

COLS = ['launch__grid_size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__cycles_active.avg', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active']
FAMILIES = {'conv_wgrad_tcgen05': r'wgrad3x3_halo', 'conv_halo': r'conv3x3_halo', 'conv_row': r'conv3x3_row',
            'bn_bwd_reduce': r'bn_bwd_reduce', 'bn_bwd_apply': r'bn_bwd_apply'}
UNIT_SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1.0, 'us': 1e3, 'usecond': 1e3, 'ms': 1e6,
              'msecond': 1e6, 'nsecond': 1.0}


def main():
    raw, out_csv, out_json = sys.argv[1:4]
    note = sys.argv[4] if len(sys.argv) > 4 else raw
    rows = [r for r in csv.reader(open(raw)) if r and not r[0].startswith('==')]
    header, units, data = rows[0], rows[1], rows[2:]
    idx = {n: i for i, n in enumerate(header)}
    kn = idx['Kernel Name']

    def val(r, name):
        i = idx.get(name)
        if i is None or r[i] in ('', 'n/a'):
            return None
        v = float(r[i].replace(',', ''))
        return v * UNIT_SCALE.get(units[i], 1.0)

    with open(out_csv, 'w') as f:
        f.write('# %s\n' % note)
        f.write('Kernel Name,' + ','.join(COLS) + '\n')
        for r in data:
            f.write('"%s",' % r[kn] + ','.join('' if val(r, c) is None else '%g' % val(r, c) for c in COLS) + '\n')
    fam = {}
    for name, pat in FAMILIES.items():
        sel = [r for r in data if re.search(pat, r[kn])]
        if not sel:
            continue
        traffic = [(val(r, 'dram__bytes_read.sum') or 0.0) + (val(r, 'dram__bytes_write.sum') or 0.0) for r in sel]
        pipe = [val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') or 0.0 for r in sel]
        fam[name] = {'launches_captured': len(sel), 'dram_mbytes_per_launch': round(sum(traffic) / len(sel) / 1e6, 1),
                     'tensor_pipe_active_pct': round(sum(pipe) / len(pipe), 1), 'source': note}
    json.dump(fam, open(out_json, 'w'), indent=1)
    print(json.dumps(fam, indent=1))


if __name__ == '__main__':
    main()
