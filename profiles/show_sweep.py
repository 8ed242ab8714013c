"""Pretty-prints quick_step.py JSON lines from stdin (one row per run)."""
import json
import sys

for l in sys.stdin:
    if not l.startswith('{'):
        if 'gpurun' in l:
            print(l.strip()[:200])
        continue
    d = json.loads(l)
    c = d['classes']
    print('%-48s step %.4f prof %.4f | bnf %.4f bnb %.4f fwd %.4f dgrad %.4f wgrad %.4f c1 %.4f head %.4f adam %.4f' % (
        d['label'][:48], d['ms_per_step'], d['profile_sum_ms'], c.get('bn_forward', 0), c.get('bn_backward', 0),
        c.get('conv_fwd_tcgen05', 0), c.get('conv_dgrad_tcgen05', 0), c.get('conv_wgrad_tcgen05', 0),
        c.get('conv_cuda_core', 0), c.get('head_loss', 0), c.get('adam_pack', 0)))
