#!/usr/bin/env python
"""VBx x-vector path (BASELINE configs[3]) on synthetic audio: K4 features + K5 ResNet101, timed with CUDA events.
   python tests/tools/vbx_profile.py [minutes]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import synth_audio                                            # noqa: E402
from inaspeechsegmenter_b200 import engine, vbx_segmenter as vb            # noqa: E402
from oracle import vbx_oracle as vx                                        # noqa: E402

minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
ctx = engine.Context(0)
fe = vb.VbxFrontEnd(ctx)
ext = vb.B200BackendExtractor(state_dict=vx.synthetic_resnet101_state(seed=5), ctx=ctx)
pcm = torch.from_numpy(synth_audio(60 * minutes, seed=4)).cuda()
fe(pcm)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
e[0].record()
fea = fe(pcm)
e[1].record()
plan = vb.window_plan(fea.shape[0])
starts = [s for s, n, tail in plan if not tail]
ext.embed_windows(fea, starts[:256], 144)
torch.cuda.synchronize()
e[2].record()
emb = ext.embed_windows(fea, starts, 144)
e[3].record()
torch.cuda.synchronize()
t_fea, t_net = e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])
M = fea.shape[0]
print('K4: %d frames in %.2f ms  (%.1f GB/s algorithmic @ 1856 B/frame)' % (M, t_fea, M * 1856 / t_fea / 1e6))
print('K5: %d windows in %.1f ms = %.0f windows/s = %.0fx real time, %.1f TFLOP/s useful'
      % (len(starts), t_net, len(starts) / t_net * 1e3, len(starts) * 0.24 / (t_net * 1e-3), ext.flops_per_window * len(starts) / t_net / 1e9))
