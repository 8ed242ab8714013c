#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (BASELINE.json: "U-Net train slices/s @256^2").

  python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle restatement of the reference's
                                                          # TF2 path on all host threads (TF is not installable)

A step = one fit step of the 4-level / 32-filter U-Net on a batch of 32 synthetic 256x256 SAX slices per GPU:
forward (batch-stat BN, dropout on) + MSE heat-map loss + full backward + gradient all-reduce (N>1) + Adam.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIG = {'DIM': [256, 256], 'DEPTH': 4, 'FILTERS': 32, 'IMG_CHANNELS': 1, 'MASK_CLASSES': 2,
          'BATCH_NORMALISATION': True, 'BN_FIRST': False, 'ACTIVATION': 'relu', 'PAD': 'same', 'KERNEL_INIT': 'he_normal',
          'M_POOL': [2, 2], 'F_SIZE': [3, 3], 'DROPOUT_MIN': 0.3, 'DROPOUT_MAX': 0.5, 'OPTIMIZER': 'adam',
          'LEARNING_RATE': 1e-4, 'SEED': 42, 'PRECISION': 'bf16'}
BATCH_PER_GPU = 32
WORKLOAD = 'C2: U-Net d4 f32 fwd+bwd+MSE+Adam, batch 32/GPU, 256x256x1 -> 2 RVIP heat maps, bf16 storage / fp32 accumulate'


def layer_table(cfg=CONFIG):
    """The 3x3 conv layers of the reference graph (Unets.py:786-836) with what follows them:
    (name, kind, cin, cout, h, w, post) -- kind 'block' = conv_layer_fn (BatchNorm), 'up' = decoder up-conv (ReLU only)."""
    H, W = cfg['DIM']
    f, cin, d = cfg['FILTERS'], cfg['IMG_CHANNELS'], cfg['DEPTH']
    out = []
    h, w = H, W
    for l in range(d):
        out += [('enc%d.conv_a' % l, 'block', cin, f, h, w, 'dropout'), ('enc%d.conv_b' % l, 'block', f, f, h, w, 'pool')]
        cin, f, h, w = f, f * 2, h // 2, w // 2
    out += [('mid.conv_a', 'block', cin, f, h, w, 'dropout'), ('mid.conv_b', 'block', f, f, h, w, 'upsample')]
    low = f
    for l in range(d):
        f //= 2
        h, w = h * 2, w * 2
        out += [('dec%d.upconv' % l, 'up', low, f, h, w, 'none'), ('dec%d.conv_a' % l, 'block', 2 * f, f, h, w, 'dropout'),
                ('dec%d.conv_b' % l, 'block', f, f, h, w, 'upsample' if l < d - 1 else 'head')]
        low = f
    return out


def bench_config(world):
    """`config` of the JSON line: identical in the B200 arm and the CPU reference arm."""
    return {'workload': WORKLOAD, 'global_batch': world * BATCH_PER_GPU, 'parallelism': 'dp%d' % world}


def conv_flops_per_slice(cfg=CONFIG):
    """Algorithmic conv FLOPs (2*MACs) per slice of the REFERENCE graph (3x3 convs on the up-sampled tensors): forward,
    and forward+dgrad+wgrad (first layer has no dgrad).  `executed_*` = what the tensor pipe runs: the phase-decomposed
    up-convolutions execute 16/36 of their forward / dgrad MMAs and 8/20 of their weight-gradient MMAs."""
    layers = layer_table(cfg)
    fl = [2 * 9 * ci * co * hh * ww for _, _, ci, co, hh, ww, _ in layers]
    fwd, first = sum(fl), fl[0]
    up = sum(v for v, l in zip(fl, layers) if l[1] == 'up')
    tc_fwd = fwd - first                  # tensor-core share (everything but the Cin=1 layer)
    return dict(fwd=fwd, train=3 * fwd - first, tc_fwd=tc_fwd, tc_dgrad=tc_fwd, tc_wgrad=tc_fwd,
                executed_fwd=tc_fwd - up + up * 16 / 36, executed_dgrad=tc_fwd - up + up * 16 / 36,
                executed_wgrad=tc_fwd - up + up * 8 / 20)


def algorithmic_bytes_per_slice(cfg=CONFIG, n_params=8635842, head_fold=True):
    """Minimum HBM bytes per slice of every memory-bound kernel class (bf16 activations, fp32 image / heat maps): each
    tensor a pass NEEDS is counted once -- BatchNorm forward = read relu(conv), write the block output (+ the pooled
    quarter); BatchNorm backward = read dy, read relu(conv), write dz (3 passes; the separate statistics pass the kernels
    run today is NOT algorithmic); the convs = read every input once (the up-conv its LOW-resolution input), write the output."""
    layers = layer_table(cfg)
    nc = cfg['MASK_CLASSES']
    b = dict(conv=0.0, bn_forward=0.0, bn_backward=0.0, conv_cuda_core=0.0, head_loss=0.0)
    for i, (name, kind, ci, co, h, w, post) in enumerate(layers):
        px = h * w
        if i == 0:
            b['conv_cuda_core'] += px * (4 * ci + 2 * co) + px * (2 * co + 4 * ci)       # forward; weight gradient
        else:
            b['conv'] += (px // 4 if kind == 'up' else px) * ci * 2 + px * co * 2
        if kind == 'block':
            if not (post == 'head' and head_fold):
                b['bn_forward'] += 2 * px * co * 2 + (px // 4 * co * 2 if post == 'pool' else 0)
            b['bn_backward'] += 3 * px * co * 2 + (px // 4 * co * 2 if post == 'pool' else 0)
        else:
            b['bn_backward'] += 3 * px * co * 2                                          # ReLU backward of the up-conv
    H, W = cfg['DIM']
    b['head_loss'] = H * W * (2 * 2 * cfg['FILTERS'] + 2 * 4 * nc)                      # read a, write dy; target, heat
    return b, 7 * 4 * n_params + 2 * 2 * 2 * n_params     # per STEP: Adam streams p, g, m, v in / p, m, v out + operand copies


class ClockSampler:
    """Samples SM clock, power and throttle reasons DURING the timed region (B200_PROFILING.md recipe) through NVML
    every few milliseconds; falls back to an `nvidia-smi -lms` child process when pynvml is unavailable."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0, period_s=0.004):
        self.rows, self.proc, self.gpu, self.period = [], None, gpu_index, period_s
        self.samples, self._stop, self._thread, self.nvml = [], False, None, None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
            idx = self.gpu
            if vis and all(v.strip().isdigit() for v in vis.split(',')):
                idx = int(vis.split(',')[self.gpu])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop:
            try:
                mhz = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    reasons = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    reasons = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                watts = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                self.samples.append((time.perf_counter(), mhz, reasons, watts))
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def mark(self):
        return time.perf_counter()

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken between the wall-clock marks t0 and t1 (all samples if none fall inside)."""
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
        if self.nvml is not None:
            inside = [s for s in self.samples if (t0 is None or s[0] >= t0) and (t1 is None or s[0] <= t1)] or self.samples
            sm = sorted(s[1] for s in inside)
            bits = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40}
            reasons = sorted({k for s in inside for k, b in bits.items() if s[2] & b})
            return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_min_mhz': sm[0] if sm else None,
                    'sm_max_mhz': self.max_mhz, 'reasons': reasons, 'samples': len(inside),
                    'power_w_max': round(max((s[3] for s in inside), default=0.0), 1), 'source': 'nvml'}
        rows = [r for t, r in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1)] or [r for _, r in self.rows]
        sm = sorted(int(float(r[1])) for r in rows if len(r) > 8 and r[1].replace('.', '').isdigit())
        mx = [int(float(r[2])) for r in rows if len(r) > 8 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in rows if len(r) > 8 for n, v in zip(names, r[5:9]) if v.lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(rows), 'source': 'nvidia-smi'}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        PEAKS['burst_tflops'] = d.get('bf16_tflops', 1646.8)
        return d.get('bf16_tflops_sustained', 1377.3), d.get('hbm_gbs', 6548.2), 'measured (MEASURED_PEAKS.json)'
    return 1400.0, 6650.0, 'fallback (B200_PROFILING.md)'


PEAKS = {'burst_tflops': 1646.8}


# ------------------------------------------------------------------------------------------ CPU arm
def oracle_train_rate(n_slices=4, steps=3, warmup=1, threads=None):
    """Times the CPU oracle (torch CPU ops, fp32) on the same net / synthetic data: slices per second."""
    import torch
    from cmr_landmark_detection_b200 import synth
    from oracle import unet_ref as R
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core it can get
    torch.set_num_threads(threads or os.cpu_count() or 1)
    cfg = R.cfg_from_config(CONFIG)
    ws = R.init_weights(cfg, seed=1234)
    x, y = synth.make_batch(n_slices, 256, 256, seed=42)
    # dropout on in the GPU arm; the oracle needs explicit masks -> Bernoulli masks of the same rates
    import numpy as np
    rng = np.random.default_rng(0)
    masks, f, h = {}, CONFIG['FILTERS'], 256
    for l in range(4):
        masks['enc%d' % l] = (rng.random((n_slices, h, h, f)) >= cfg.dropouts[l]).astype(np.float32)
        masks['dec%d' % (3 - l)] = (rng.random((n_slices, h, h, f)) >= cfg.dropouts[l]).astype(np.float32)
        f, h = f * 2, h // 2
    masks['mid'] = (rng.random((n_slices, h, h, f)) >= cfg.drop_mid).astype(np.float32)
    opt = R.Adam(lr=1e-4)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = R.train_grads(cfg, ws, x, y, dropout_masks=masks)
        ws = opt.step(R.apply_new_stats(cfg, ws, out['new_stats']), out['grads'])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return n_slices * len(times) / sum(times), torch.get_num_threads(), sum(times) / len(times)


def run_reference(args):
    """CPU arm: the oracle restatement of the reference's TF2 path (TensorFlow is not installable here) on every host
    thread, same metric / config / warm-up count as the B200 arm; every step is a bounded 4-slice sample of the
    32-slice batch so that W + K steps end within minutes."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_slices = 4
    W = max(args.warmup, 3)
    rate, cores, spt = oracle_train_rate(n_slices=n_slices, steps=args.steps, warmup=W)
    line = {'impl': 'reference', 'metric': 'unet_train_slices_per_s_256', 'value': rate, 'unit': 'slices/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': W, 'ms_per_step': spt * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': bench_config(args.gpus),
            'notes': {'what': 'CPU restatement (torch/oneDNN, fp32) of the reference TF2 path; TF is not installable here',
                      'sample': 'each step = %d slices of the 32-slice batch (bounded sample)' % n_slices},
            'cpu_baseline': {'value': rate, 'unit': 'slices/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d-slice fwd+bwd+Adam step x %d (after %d warm-up steps)' % (n_slices, args.steps, W)},
            'e2e': {'value': rate, 'unit': 'slices/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import numpy as np
    import torch
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.models.Unets import create_unet
    from cmr_landmark_detection_b200.runtime import dist as rdist
    rank, local, world = rdist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    model = create_unet(dict(CONFIG))
    B = BATCH_PER_GPU
    # two different synthetic batches alternate so no step sees cache-warm inputs; the per-step working set
    # (~4.4 GB of activations + gradients) is far larger than the 126 MB L2 anyway
    data = [synth.make_batch(B, 256, 256, seed=42 + 7 * rank + i) for i in range(2)]
    dev_data = [(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)) for x, y in data]
    K, Wm = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(Wm):
        model.train_step_device(*dev_data[i % 2])
    barrier()
    t_mark0 = sampler.mark()
    l0 = model.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = model.train_step_device(*dev_data[i % 2])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = model.launch_count() - l0
    clocks = sampler.stop(t_mark0, sampler.mark()) if rank == 0 else None
    ms = model.dp.max_float(ms)
    value = world * B * K / (ms * 1e-3)
    final_loss = float(loss.item())

    # ---- end to end through the public API: model.fit() over a keras.utils.Sequence-like object that hands out
    # HOST numpy batches.  Every step stages (x, y) in pinned memory, copies them to the device and reads the loss
    # back, all inside the timed region; fit() pipelines batch i+1's staging / H2D behind step i's kernels.
    class _Seq:
        rvip_per_replica = True      # weak scaling: every rank's generator hands out that rank's OWN 32-slice batches

        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            return data[i % 2]

        def on_epoch_end(self):
            pass

    model.fit(_Seq(3), epochs=1, verbose=0)
    barrier()
    e0.record()
    hist = model.fit(_Seq(K), epochs=1, verbose=0)
    e1.record()
    barrier()
    ms_e2e = model.dp.max_float(e0.elapsed_time(e1))
    e2e = world * B * K / (ms_e2e * 1e-3)
    h2d = int(data[0][0].nbytes + data[0][1].nbytes)
    assert np.isfinite(hist.history['loss'][-1])

    # ---- per-kernel-class device time (CUDA events around every launch group) for the roofline
    model.profile(B, True, True)
    for i in range(K):
        model.train_step_device(*dev_data[i % 2])
    torch.cuda.synchronize()
    prof = model.profile_read(B, True)
    model.profile(B, True, False)

    line = None
    if rank == 0:
        peak_tf, peak_gbs, peak_src = peaks()
        burst_tf = PEAKS['burst_tflops']
        fl = conv_flops_per_slice()
        by, adam_bytes = algorithmic_bytes_per_slice(n_params=model.n_params)
        per_step = {k: v[0] / K for k, v in prof.items()}
        tot = sum(per_step.values())
        # ALGORITHMIC work of every kernel class per step: conv FLOPs of the reference graph for the tensor-core classes,
        # minimum HBM bytes (algorithmic_bytes_per_slice) for the streaming classes
        cls_flops = {'conv_fwd_tcgen05': fl['tc_fwd'] * B, 'conv_dgrad_tcgen05': fl['tc_dgrad'] * B,
                     'conv_wgrad_tcgen05': fl['tc_wgrad'] * B}
        cls_exec = {'conv_fwd_tcgen05': fl['executed_fwd'] * B, 'conv_dgrad_tcgen05': fl['executed_dgrad'] * B,
                    'conv_wgrad_tcgen05': fl['executed_wgrad'] * B}
        cls_bytes = {'bn_forward': by['bn_forward'] * B, 'bn_backward': by['bn_backward'] * B,
                     'conv_cuda_core': by['conv_cuda_core'] * B, 'head_loss': by['head_loss'] * B, 'adam_pack': adam_bytes}
        kern = {}
        for k, t in per_step.items():
            if t <= 0:
                continue
            e = {'ms_per_step': round(t, 4), 'share': round(t / tot, 4), 'launches_per_step': prof[k][1] // K}
            if k in cls_flops:
                tf = cls_flops[k] / (t * 1e-3) / 1e12
                e.update(bound='tensor', achieved=round(tf, 1), unit='TFLOP/s', frac=round(tf / peak_tf, 4),
                         frac_of_burst=round(tf / burst_tf, 4),
                         executed_tflops=round(cls_exec[k] / (t * 1e-3) / 1e12, 1),
                         executed_over_algorithmic=round(cls_exec[k] / cls_flops[k], 4))
            elif k in cls_bytes:
                gbs = cls_bytes[k] / (t * 1e-3) / 1e9
                e.update(bound='hbm', achieved=round(gbs, 1), unit='GB/s', frac=round(gbs / peak_gbs, 4),
                         algorithmic_mbytes=round(cls_bytes[k] / 1e6, 1))
            kern[k] = e
        # the dominant class = the one with the largest share of the step (whatever roof binds it)
        dom = max((k for k in kern if 'bound' in kern[k]), key=lambda k: per_step[k])
        conv_t = sum(per_step.get(k, 0.0) for k in cls_flops)
        conv_tf = sum(cls_flops.values()) / (conv_t * 1e-3) / 1e12
        # DRAM traffic per step of the dominant class from the committed ncu --set full capture of THIS code
        # (profiles/capture.sh full -> profiles/r2_traffic.json); null when the capture does not cover it
        traffic, traffic_src, ncu_extra = None, None, None
        tpath = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
        if os.path.exists(tpath):
            tj = json.load(open(tpath)).get(dom)
            if tj:
                traffic = int(tj['dram_mbytes_per_step'] * 1e6)
                traffic_src = tj['source']
                ncu_extra = {k: v for k, v in tj.items() if k not in ('dram_mbytes_per_step', 'source')}
        step_tf = fl['train'] * B / (ms / K * 1e-3) / 1e12
        roof = {'bound': kern[dom]['bound'], 'kernel': dom, 'achieved': kern[dom]['achieved'],
                'peak': peak_tf if kern[dom]['bound'] == 'tensor' else peak_gbs, 'unit': kern[dom]['unit'],
                'frac': kern[dom]['frac'], 'traffic': traffic, 'traffic_source': traffic_src, 'ncu': ncu_extra,
                'peak_source': peak_src, 'peaks': {'hbm_gbs': peak_gbs, 'bf16_tflops_sustained': peak_tf,
                                                   'bf16_tflops_burst': burst_tf},
                'all_conv_tcgen05_tflops': round(conv_tf, 1), 'all_conv_frac': round(conv_tf / peak_tf, 4),
                'step_tflops_algorithmic': round(step_tf, 1),
                'step_tensor_frac': round(step_tf / peak_tf, 4), 'step_tensor_frac_of_burst': round(step_tf / burst_tf, 4),
                'note': 'achieved = ALGORITHMIC work / device time per class (CUDA events around every launch group, '
                        'weight gradients not overlapped in this profiling pass); the phase-decomposed up-convolutions '
                        'execute fewer MMAs than the reference graph has (executed_over_algorithmic)',
                'kernels': kern}
        # CPU baseline: rank 0 at N = 1 only (under torchrun OMP_NUM_THREADS=1 and N ranks share the host cores)
        cpu = None
        if world == 1:
            cpu_rate, cores, spt = oracle_train_rate(n_slices=4, steps=2, warmup=1)
            cpu = {'value': round(cpu_rate, 3), 'unit': 'slices/s', 'cores': cores, 'kind': 'port',
                   'sample': '4-slice fwd+bwd+Adam step x 2 (oracle, torch CPU fp32), %.2f s/step' % spt}
        line = {'metric': 'unet_train_slices_per_s_256', 'value': round(value, 1), 'unit': 'slices/s', 'n_gpus': world,
                'steps': K, 'warmup': Wm, 'ms_per_step': round(ms / K, 4), 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
                'config': bench_config(world),
                'notes': {'l2': 'per-step working set ~4.4 GB >> 126 MB L2; two input batches alternate',
                          'final_loss': final_loss, 'gflop_per_slice_train': round(fl['train'] / 1e9, 2)},
                'clocks': clocks, 'gpu_launches': int(launches),
                'e2e': {'value': round(e2e, 1), 'unit': 'slices/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 8,
                        'ms_per_step': round(ms_e2e / K, 4), 'api': 'model.fit(Sequence of host batches)',
                        'host_thread_ms_per_step': getattr(model, 'last_fit_timing', None)},
                'roofline': roof,
                'cpu_baseline': cpu}
        # secondary metrics build further models on this rank only: single-process runs only (a data-parallel model
        # issues collectives the other ranks would never match)
        if not args.no_extra and world == 1:
            line['extra'] = extra_measurements(model, dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def extra_measurements(model, dev):
    """Secondary BASELINE metrics: volume inference (C4: 16 x 256 x 256) with fused extraction, and the
    extraction kernel alone against the HBM roofline."""
    import torch
    from cmr_landmark_detection_b200 import synth
    from cmr_landmark_detection_b200.extract import extract_device
    out = {}
    x, _ = synth.make_batch(16, 256, 256, seed=3)
    xd = torch.from_numpy(x).to(dev)
    for _ in range(3):
        extract_device(model.predict_device(xd))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        r = extract_device(model.predict_device(xd))
    e1.record()
    torch.cuda.synchronize()
    out['infer_vols_per_s'] = round(n / (e0.elapsed_time(e1) * 1e-3), 1)
    # the same through the public API with HOST buffers: model.predict(ndarray) -> ndarray (pinned staging, H2D of the
    # slices, D2H of the heat maps inside the timed region), at one volume per call and at pred_fold's batch size 1
    # (predict_model.py:89: one forward per slice)
    for _ in range(2):
        model.predict(x, batch_size=16)
        model.predict(x, batch_size=1)
    t0 = time.perf_counter()
    for _ in range(n):
        hp = model.predict(x, batch_size=16)
    t1 = time.perf_counter()
    for _ in range(3):
        model.predict(x, batch_size=1)
    t2 = time.perf_counter()
    out['infer_e2e'] = {'api': 'model.predict(host ndarray [16,256,256,1]) -> host ndarray', 'vols_per_s': round(n / (t1 - t0), 1),
                        'h2d_bytes_per_vol': int(x.nbytes), 'd2h_bytes_per_vol': int(hp.nbytes),
                        'batch1_slices_per_s': round(3 * 16 / (t2 - t1), 1), 'batch1_ms_per_slice': round((t2 - t1) / 48 * 1e3, 3)}
    heat = torch.from_numpy(synth.make_volume_heat(16 * 64, 256, 256, seed=1)[:, :, :, :]).to(dev)   # 537 MB > L2
    for _ in range(3):
        extract_device(heat)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        extract_device(heat)
    e1.record()
    torch.cuda.synchronize()
    gbs = heat.numel() * 4 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
    out['extract_gbs'] = round(gbs, 1)
    out['extract_frac_of_hbm_peak'] = round(gbs / peaks()[1], 4)
    del heat
    # throughput-oriented inference: 8 volumes per call
    x8, _ = synth.make_batch(128, 256, 256, seed=4)
    x8d = torch.from_numpy(x8).to(dev)
    for _ in range(2):
        extract_device(model.predict_device(x8d))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        extract_device(model.predict_device(x8d))
    e1.record()
    torch.cuda.synchronize()
    out['infer_vols_per_s_8_per_call'] = round(5 * 8 / (e0.elapsed_time(e1) * 1e-3), 1)
    del x8d
    # BASELINE config C5 (tensor-core stress): 5 levels, 64 base filters, 512 x 512, batch 8
    try:
        from cmr_landmark_detection_b200.models.Unets import create_unet
        c5 = dict(CONFIG, DIM=[512, 512], DEPTH=5, FILTERS=64, DATA_PARALLEL=False)
        m5 = create_unet(c5)
        x5, y5 = synth.make_batch(8, 512, 512, seed=5)
        x5d, y5d = torch.from_numpy(x5).to(dev), torch.from_numpy(y5).to(dev)
        for _ in range(3):
            m5.train_step_device(x5d, y5d)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            m5.train_step_device(x5d, y5d)
        e1.record()
        torch.cuda.synchronize()
        ms5 = e0.elapsed_time(e1) / 5
        fl5 = conv_flops_per_slice(c5)['train'] * 8
        out['c5_train'] = {'workload': 'C5: U-Net d5 f64 fwd+bwd+MSE+Adam, batch 8, 512x512', 'ms_per_step': round(ms5, 3),
                           'slices_per_s': round(8 / ms5 * 1e3, 1), 'conv_tflops': round(fl5 / ms5 / 1e9, 1),
                           'frac_of_sustained_bf16_peak': round(fl5 / ms5 / 1e9 / peaks()[0], 4)}
        del m5
    except Exception as e:          # secondary metric: never fail the bench line over it
        out['c5_train'] = {'error': str(e)[:200]}
    # the reference's own training resolution (DIM [224, 224], Train_tests.ipynb): 224 is off the 128-pixel row tiles and
    # only its 112-pixel level fits the 16-pixel halo blocks of the phase-decomposed up-conv
    try:
        c224 = dict(CONFIG, DIM=[224, 224], DATA_PARALLEL=False)
        m2 = create_unet(c224)
        x2, y2 = synth.make_batch(32, 224, 224, seed=6)
        x2d, y2d = torch.from_numpy(x2).to(dev), torch.from_numpy(y2).to(dev)
        for _ in range(3):
            m2.train_step_device(x2d, y2d)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            m2.train_step_device(x2d, y2d)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 10
        out['train_224'] = {'workload': 'same net and batch at the reference\'s DIM 224x224', 'ms_per_step': round(ms2, 3),
                            'slices_per_s': round(32 / ms2 * 1e3, 1),
                            'conv_tflops': round(conv_flops_per_slice(c224)['train'] * 32 / ms2 / 1e9, 1)}
        del m2
    except Exception as e:
        out['train_224'] = {'error': str(e)[:200]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-extra', action='store_true', help='skip the secondary BASELINE metrics (volume inference '
                    'vols/s with fused extraction, extraction GB/s)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
