"""Counts the Blackwell tensor-core / TMA / TMEM SASS mnemonics per kernel of librvip_b200.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA tensor load / store), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit),
SYNCS (mbarrier).  usage: python profiles/sass_summary.py > profiles/<round>_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'cmr_landmark_detection_b200', 'librvip_b200.so')
MNEMONICS = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'UTCBAR', 'SYNCS', 'HMMA', 'RED', 'ATOM']


def main():
    out = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m:
            op = m.group(1)
            per[cur]['_total'] += 1
            for mn in MNEMONICS:
                if op.startswith(mn):
                    per[cur][mn] += 1
    demangled = subprocess.run(['c++filt'], input='\n'.join(per), capture_output=True, text=True).stdout.splitlines()
    print('# %s: SASS mnemonic counts per kernel (sm_100a)' % os.path.basename(SO))
    print('%-110s %8s ' % ('kernel', 'instrs') + ' '.join('%8s' % m for m in MNEMONICS))
    tot = collections.Counter()
    for (name, c), dn in zip(per.items(), demangled):
        dn = re.sub(r'\(.*$', '', dn).replace('rvip::', '')
        if not any(c[m] for m in MNEMONICS[:6]):
            continue
        print('%-110s %8d ' % (dn[:110], c['_total']) + ' '.join('%8d' % c[m] for m in MNEMONICS))
        tot.update(c)
    print('%-110s %8d ' % ('TOTAL (kernels listed)', tot['_total']) + ' '.join('%8d' % tot[m] for m in MNEMONICS))


if __name__ == '__main__':
    sys.exit(main())
