#!/usr/bin/env python
"""K1 (sidekit features) on the GPU: accuracy against the numpy oracle (pinned bit-for-bit to sidekit_mfcc.py) on a
60 s synthetic clip + the committed media, and throughput on N hours of synthetic int16 PCM (CUDA events).
   python tests/tools/k1_check.py [hours=10]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench                                                               # noqa: E402
from conftest import synth_audio                                           # noqa: E402
from inaspeechsegmenter_b200 import _lib, io as iss_io                     # noqa: E402
from inaspeechsegmenter_b200.segmenter import feats_from_signal            # noqa: E402
from oracle import sidekit_oracle as so                                    # noqa: E402

hours = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
clips = {'synth60': synth_audio(60.0, seed=11)}
media = os.path.join(ROOT, 'tests', 'golden', 'media')
for f in ('musanmix.wav', 'lamartine.wav'):
    clips[f] = iss_io.media2sig16kmono(os.path.join(media, f), ffmpeg=None, dtype='float32')
for prec, name in ((_lib.FFT_FP64, 'fp64'), (_lib.FFT_FP32, 'fp32')):
    for cname, sig in clips.items():
        sig = np.asarray(sig, dtype=np.float32)
        mspec, loge, _ = feats_from_signal(torch.from_numpy(sig).cuda(), 0, prec, 'main')
        ref_m, ref_e = so.logmel_loge(sig)
        m, e = mspec.cpu().numpy(), loge.cpu().numpy()
        fin = np.isfinite(ref_m)
        print('%s %-14s mspec max abs err %.3e  bit-identical %.1f %%  inf pattern equal %s ; loge max abs err %.3e'
              % (name, cname, np.abs(m[fin] - ref_m[fin]).max(), 100.0 * np.mean(m[fin] == ref_m[fin]),
                 bool(np.array_equal(np.isfinite(m), fin)), np.abs(e[np.isfinite(ref_e)] - ref_e[np.isfinite(ref_e)]).max()), flush=True)
n = int(hours * 3600 * 16000)
pcm = bench.synth_range(torch, 0, n, torch.device('cuda:0'))
for prec, name in ((_lib.FFT_FP64, 'fp64'), (_lib.FFT_FP32, 'fp32')):
    feats_from_signal(pcm, 0, prec, 'main')
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(3):
        mspec, loge, _ = feats_from_signal(pcm, 0, prec, 'main')
    ev[1].record(); ev[1].synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 3
    L = loge.numel()
    print('%s: %d frames (%g h) in %.2f ms = %.1f GB/s algorithmic (420 B/frame), %.3f ms per audio-hour' % (name, L, hours, ms, 420.0 * L / ms / 1e6, ms / hours), flush=True)
