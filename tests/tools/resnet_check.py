#!/usr/bin/env python
"""ResNet101 (K5) accuracy/throughput per GEMM engine against the committed resnet.py golden."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from inaspeechsegmenter_b200 import _lib, engine, vbx_segmenter as vb     # noqa: E402
from oracle import vbx_oracle as vx                                        # noqa: E402

g = np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_golden.npz'))
lib = _lib.load()
ctx = engine.Context(0)
sd = vx.synthetic_resnet101_state(seed=5)
ext = vb.B200BackendExtractor(state_dict=sd, ctx=ctx)
x, y = g['resnet_x'], g['resnet_y']
only = (int(sys.argv[1]),) if len(sys.argv) > 1 else (0, 2, 3)          # e.g. `resnet_check.py 3`: the default engine only, longer timing
nw = 1024 if len(sys.argv) > 1 else 256
fea = torch.randn(24 * nw + 144, 64, device='cuda')
starts = np.arange(nw) * 24
for mode in only:
    lib.iss_set_gemm_mode(mode)
    errs = [np.abs(ext.get_embedding(x[i].T) - y[i]).max() / np.abs(y).max() for i in range(len(x))]
    ext.embed_windows(fea, starts, 144)
    torch.cuda.synchronize()
    reps = 3 if len(sys.argv) > 1 else 1
    t0 = time.perf_counter()
    for _ in range(reps):
        ext.embed_windows(fea, starts, 144)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print('mode %d: max rel err vs resnet.py golden = %.3e ; %d windows in %.1f ms = %.1f TFLOP/s, %.0f windows/s'
          % (mode, max(errs), len(starts), dt * 1e3, ext.flops_per_window * len(starts) / dt / 1e12, len(starts) / dt), flush=True)
