#!/usr/bin/env python
"""bench.py -- audio-hours/sec of the inaSpeechSegmenter hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One step = one pass of the whole hot path (K1 log-mel/energy -> energy Viterbi
-> smn CNN -> Viterbi -> gender CNN -> Viterbi -> segment list on the host) over
one batch of synthetic 16 kHz mono int16 audio: BASELINE.json configs[1]
("smn+gender on 10 h synthetic, 1xB200").  With N > 1 ranks (torchrun) ONE recording
of N x 10 h is time-sharded over the ranks (inaspeechsegmenter_b200/shard.py:
34-frame halo, NCCL all-gather of loge and of the CNN posteriors, BASELINE
configs[4]) => weak scaling (per-GPU work fixed); the timed region is bracketed
by barrier + synchronize and the max over ranks is reported.

`--impl reference` times the reference's CPU path (the numpy/torch-CPU oracle
port: TensorFlow and the .hdf5 networks are not installable here) on the host
cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'audio-hours/sec segmented (16 kHz mono)'
UNIT = 'audio-hours/s'
SR = 16000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--hours', type=float, default=10.0, help='audio hours per GPU per step')
    ap.add_argument('--fft', default='fp64', choices=['fp64', 'fp32'])
    ap.add_argument('--cpu-sample-sec', type=float, default=60.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


# ----------------------------------------------------------------------------- synthetic audio
def synth_block_torch(torch, block, seconds, device):
    """One block of the SURVEY 8(d) generator on the device: spans of exact
    silence, white noise, a harmonic 'speech-like' source with 4 Hz AM and
    multi-tone 'music' (5-30 s, plus a few spans < 0.68 s), quantised to int16."""
    g = torch.Generator(device='cpu')
    g.manual_seed(20260922 + block)
    n = int(seconds * SR)
    out = torch.zeros(n, dtype=torch.float32, device=device)
    gd = torch.Generator(device=device)
    gd.manual_seed(977 * (20260922 + block) + 1)
    pos = 0
    while pos < n:
        u = torch.rand(8, generator=g)
        dur = int((0.3 + 0.35 * u[0].item()) * SR) if u[1].item() < 0.08 else int((5 + 25 * u[0].item()) * SR)
        kind = int(u[2].item() * 4)
        end = min(n, pos + dur)
        m = end - pos
        t = torch.arange(m, device=device, dtype=torch.float32) / SR
        if kind == 1:
            out[pos:end] = torch.randn(m, generator=gd, device=device) * (1e-3 + 0.3 * u[3].item())
        elif kind == 2:
            f0 = 100 + 150 * u[3].item()
            s = torch.zeros(m, device=device)
            for h in range(1, 12):
                s += torch.sin(2 * np.pi * f0 * h * t + 6.28 * u[4].item() * h) / h
            out[pos:end] = 0.08 * s * (0.6 + 0.4 * torch.sin(2 * np.pi * 4 * t)) + torch.randn(m, generator=gd, device=device) * 2e-3
        elif kind == 3:
            s = torch.zeros(m, device=device)
            for q in range(5):
                s += torch.sin(2 * np.pi * (200 + 2800 * ((u[3].item() * (q + 1) * 0.618) % 1.0)) * t)
            out[pos:end] = 0.05 * s + torch.randn(m, generator=gd, device=device) * 1e-3
        pos = end
    return torch.clamp(torch.round(out * 32768), -32768, 32767).to(torch.int16)


BLOCK_SEC = 600.0


def synth_range(torch, sa, sb, device):
    """Samples [sa, sb) of the (arbitrarily long) synthetic recording made of 10-minute blocks."""
    bl = int(BLOCK_SEC * SR)
    parts = []
    for b in range(sa // bl, (sb - 1) // bl + 1):
        blk = synth_block_torch(torch, b, BLOCK_SEC, device)
        parts.append(blk[max(sa - b * bl, 0):min(sb - b * bl, bl)])
    return torch.cat(parts)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                       '-lms', '200'], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nme, v in zip(names, c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nme)
        self.f.close()
        os.unlink(self.f.name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------- CPU (reference-port) leg
def cpu_reference_pass(sample_f32, mods, threads):
    """The reference's CPU path restated (oracle): numpy front-end, numpy patch
    materialisation, torch-CPU CNNs (all host threads), C Viterbi."""
    import warnings
    from oracle import cnn_oracle, segmenter_oracle as so
    t0 = time.perf_counter()
    mspec, loge, difflen = so.media2feats(sample_f32)
    v = so.DnnSegmenterOracle(cnn_oracle.KerasLikeModel(*mods['vad'], threads=threads), **so.VAD_SMN)
    g = so.DnnSegmenterOracle(cnn_oracle.KerasLikeModel(*mods['gender'], threads=threads), **so.GENDER)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        segs = so.segment_feats(mspec, loge, difflen, 0, v, g)
    return time.perf_counter() - t0, segs


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


_CPU_THREADS = None


def cpu_threads(mods=None, sample=None):
    """Threads the CPU arm uses.  The torch-CPU convolutions of the port stop scaling well below
    128 threads (measured on the GPU box: 16 threads 0.0086, 128 threads 0.0015 audio-hours/s), so
    the best of {8, 16, 32, 64, all cores} on a 5 s slice is used -- the CPU side gets its best
    configuration.  ISS_CPU_THREADS pins it."""
    global _CPU_THREADS
    env = os.environ.get('ISS_CPU_THREADS')
    if env:
        return max(1, int(env))
    if _CPU_THREADS is None:
        cores = host_cores()
        if mods is None or sample is None or cores <= 8:
            return cores
        best, best_t = cores, None
        for n in sorted({c for c in (8, 16, 32, 64, cores) if c <= cores}):
            cpu_reference_pass(sample[:SR * 2], mods, n)
            t = cpu_reference_pass(sample[:SR * 5], mods, n)[0]
            if best_t is None or t < best_t:
                best, best_t = n, t
        _CPU_THREADS = best
    return _CPU_THREADS


def make_models():
    from inaspeechsegmenter_b200 import models
    return {'vad': models.synthetic_keras_cnn(21, 3, seed=11), 'gender': models.synthetic_keras_cnn(24, 2, seed=13)}


def cpu_sample(args):
    """Host copy of the first cpu-sample-sec seconds of rank 0's recording (generated on CPU torch)."""
    import torch
    n = int(args.cpu_sample_sec * SR)
    s16 = synth_range(torch, 0, n, 'cpu').numpy()
    return s16.astype(np.float32) / np.float32(32768)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    mods = make_models()
    sample = cpu_sample(args)
    from oracle import viterbi_oracle
    viterbi_oracle.build()
    cores = cpu_threads(mods, sample)
    for _ in range(args.warmup):
        cpu_reference_pass(sample[:SR * 10], mods, cores)
    ts = [cpu_reference_pass(sample, mods, cores)[0] for _ in range(args.steps)]
    t = float(np.mean(ts))
    val = (len(sample) / SR / 3600.0) / t
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'smn+gender on %g h synthetic 16 kHz mono (bounded sample: first %g s)' % (args.hours, args.cpu_sample_sec),
                   'networks': 'synthetic-weight stand-ins (release .hdf5 absent)'},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'host_cores': host_cores(),
                         'sample': 'first %g s of the synthetic recording; reference-numpy front-end + torch-CPU restatement of the CNNs + C Viterbi (TensorFlow absent); thread count = best of {8,16,32,64,all}' % args.cpu_sample_sec},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))
    return 0


def run_b200(args):
    import torch
    import torch.distributed as dist
    from inaspeechsegmenter_b200 import Segmenter, _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    mods = make_models()
    seg = Segmenter(vad_engine='smn', detect_gender=True, ffmpeg=None, models=mods, device=local, fft_precision=args.fft)

    from inaspeechsegmenter_b200.shard import ShardPlan, segment_signal_sharded
    total = int(args.hours * world * 3600 * SR)            # ONE recording of world x hours, time-sharded
    plan = ShardPlan(total, world)
    sa, sb = plan.sample_range(rank)
    pcm = synth_range(torch, sa, sb, dev)                               # this rank's samples (+halo), resident in HBM
    host = torch.empty(pcm.shape, dtype=torch.int16, pin_memory=True)    # pinned copy for the e2e leg
    host.copy_(pcm)
    torch.cuda.synchronize()
    audio_h = total / SR / 3600.0 / world                             # per-rank share of the recording

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        e1.synchronize()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), out

    if world == 1:
        step_dev = lambda: seg.segment_signal(pcm)          # noqa: E731  inputs already in HBM
        step_e2e = lambda: seg.segment_signal(host)         # noqa: E731  pinned host -> device inside the call
    else:
        step_dev = lambda: segment_signal_sharded(seg, pcm, total)[0]     # noqa: E731
        step_e2e = lambda: segment_signal_sharded(seg, host, total)[0]    # noqa: E731

    for _ in range(max(args.warmup, 3)):
        segs = step_dev()
    # dominant kernel = the conv/dense layer with the most FLOPs of the VAD network
    nl = lib.iss_cnn_num_layers(seg.vad.nn.handle)
    lf = [lib.iss_cnn_layer_flops(seg.vad.nn.handle, i) for i in range(nl)]
    dom = int(np.argmax(lf))
    _lib.check(lib.iss_cnn_profile(seg.vad.nn.handle, dom), 'iss_cnn_profile')

    clocks = ClockSampler(local)
    l0 = lib.iss_launch_count()
    ms, segs = timed(step_dev, args.steps)
    launches = lib.iss_launch_count() - l0
    clk = clocks.stop()
    import ctypes
    tms, nlaunch, fl = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
    _lib.check(lib.iss_cnn_profile_read(seg.vad.nn.handle, ctypes.byref(tms), ctypes.byref(nlaunch), ctypes.byref(fl)), 'profile_read')
    _lib.check(lib.iss_cnn_profile(seg.vad.nn.handle, -1), 'iss_cnn_profile')

    ms_e2e, segs2 = timed(step_e2e, args.steps)
    assert segs2 == segs
    d2h = int(seg_d2h_bytes(segs, pcm.numel()))

    value = audio_h * world * args.steps / (ms / 1e3)
    e2e = audio_h * world * args.steps / (ms_e2e / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak_tf = peaks.get('bf16_tflops_sustained') or 1400.0
    peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if 'bf16_tflops_sustained' in peaks else 'fallback 1.4 PFLOP/s sustained (of fallback)'
    ach = (fl.value / max(nlaunch.value, 1)) / (tms.value / max(nlaunch.value, 1) * 1e-3) / 1e12 if tms.value > 0 else 0.0
    mode = lib.iss_get_gemm_mode()
    gemm = {0: 'fp32 CUDA cores', 1: 'tcgen05 3xTF32 (A,B from smem)', 2: 'tcgen05 3xTF32 (A from TMEM)'}.get(mode, 'engine %d' % mode)
    traffic, traffic_src = None, None
    if mode == 2 and os.environ.get('ISS_B200_TC_SLAB', '1') != '0' and not os.environ.get('ISS_B200_TC3_CFG'):
        traffic, traffic_src = ncu_dram_bytes(os.path.join(ROOT, 'profiles', 'r01_conv_gemm_tc3_final_full.txt'), 'conv_gemm_tc3_kernel<64')
    roof = {'bound': 'tensor', 'kernel': 'conv_gemm %s (VAD layer %d: %s)' % (gemm, dom, layer_name(seg.vad.nn.lowered.descs[dom])),
            'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf, 'traffic': traffic, 'traffic_source': traffic_src,
            'peak_source': peak_src, 'launches': int(nlaunch.value), 'avg_launch_ms': tms.value / max(nlaunch.value, 1),
            'flops_per_launch': fl.value / max(nlaunch.value, 1), 'share_of_step': tms.value / ms}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32' if lib.iss_get_gemm_mode() == 0 else 'f32 (3xTF32 split on tcgen05, fp32 accumulate)', 'data': 'synthetic',
        'config': {'workload': 'smn+gender on %g h synthetic 16 kHz mono int16 per GPU (BASELINE configs[1])' % args.hours,
                   'networks': 'synthetic-weight stand-ins of the ~1.4M-parameter CNN family (release .hdf5 absent)',
                   'fft': args.fft, 'l2': 'inputs larger than L2 (%.2f GB PCM per step)' % (pcm.numel() * 2 / 1e9),
                   'parallelism': ('single GPU' if world == 1 else
                                   'one %g h recording time-sharded over %d GPUs (34-frame halo); NCCL all-gathers of loge, CNN posteriors, label tracks and (>= 3 ranks) max-plus transfer matrices of the energy Viterbi' % (args.hours * world, world)),
                   'segments': len(segs),
                   'vad_flops_per_patch': seg.vad.nn.flops_per_patch, 'gender_flops_per_patch': seg.gender.nn.flops_per_patch},
        'clocks': clk,
        'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': int(pcm.numel() * 2), 'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'roofline': roof,
    }
    if not args.no_cpu_baseline and world == 1:
        sample = cpu_sample(args)
        cores = cpu_threads(mods, sample)
        cpu_reference_pass(sample[:SR * 5], mods, cores)
        t, _ = cpu_reference_pass(sample, mods, cores)
        line['cpu_baseline'] = {'value': (len(sample) / SR / 3600.0) / t, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'host_cores': host_cores(),
                                'sample': 'first %g s of the recording: numpy front-end + torch-CPU CNN restatement + C Viterbi; thread count = best of {8,16,32,64,all}' % args.cpu_sample_sec}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def ncu_dram_bytes(path, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) of `kernel` from a committed
    `ncu --set full` summary of the same 2048-patch launch; (None, None) if the file is not there."""
    try:
        tot, hit = 0.0, False
        for line in open(path):
            if line.startswith('kernel:'):
                if hit:
                    break
                hit = kernel in line
            elif hit and ('dram__bytes_read.sum ' in line or 'dram__bytes_write.sum ' in line):
                val, unit = line.split()[-2:]
                tot += float(val) * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
        if hit and tot > 0:
            return tot, 'ncu --set full, %s (bytes per launch, cold cache)' % os.path.relpath(path, ROOT)
    except Exception:
        pass
    return None, None


def layer_name(d):
    kind = {1: 'Conv2D', 2: 'Dense', 3: 'MaxPool'}[d['kind']]
    return '%s %dx%d %d->%d' % (kind, d['kh'], d['kw'], d['cin'], d['cout'])


def seg_d2h_bytes(segs, n_samples):
    """Bytes read back per step: the three uint8 label tracks (energy [P], VAD and gender tracks over evaluated patches)."""
    L = (n_samples - 400) // 160 + 1
    return 3 * ((L + 1) // 2)


if __name__ == '__main__':
    a = parse()
    sys.exit(run_reference(a) if a.impl == 'reference' else run_b200(a))
