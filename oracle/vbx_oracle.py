"""Oracle (test infrastructure): the VBx x-vector path.

* ``get_features`` restates reference ``vbx_segmenter.py:72-89`` on top of a
  restatement of ``features_vbx.py`` (``povey_window`` :123, ``mel_fbank_mx``
  :31-59, ``add_dither`` :127, ``fbank_htk`` :62-120 with USEPOWER=True /
  ZMEANSOURCE=True / PREEMCOEF=0.97, ``cmvn_floating_kaldi`` :131-148).
  PINNED bit-for-bit against the real ``features_vbx.py`` by
  ``tests/golden/make_golden.py``.
* ``window_plan`` restates the windowing loop of ``VBxExtractor.__call__``
  (``vbx_segmenter.py:217-246``).
* ``ResNet101Oracle`` restates the architecture of ``resnet.py:48-135``
  (Bottleneck [3,4,23,3], stat-pooling, Linear) in functional torch-CPU fp32
  over a plain state_dict.  PARITY UNPINNED for the real weights
  (``final.onnx`` / ``raw_81.pth`` are release assets, absent here).
"""
import numpy as np
import torch
import torch.nn.functional as F

SR, STEP, WINLEN, FEAT_DIM, EMBED_DIM = 16000, 24, 144, 64, 256      # vbx_segmenter.py:21-25
NOVERLAP, FRAME, NFFT = 240, 400, 512


def povey_window(n=FRAME):
    return np.power(0.5 - 0.5 * np.cos(np.linspace(0, 2 * np.pi, n)), 0.85)


def mel_bank(nfft=NFFT, fs=SR, nchan=FEAT_DIM, lo=20.0, hi=7600.0):
    """mel_fbank_mx(400, 16000, NUMCHANS=64, LOFREQ=20, HIFREQ=7600, htk_bug=False): [257, 64] float64,
    unnormalised triangles on the 1127*ln(1+f/700) scale."""
    mel = lambda x: 1127. * np.log(1. + x / 700.)            # noqa: E731
    imel = lambda x: (np.exp(x / 1127.) - 1.) * 700.         # noqa: E731
    fbin = mel(np.arange(nfft / 2 + 1, dtype=float) * fs / nfft)
    cbin = np.linspace(mel(lo), mel(hi), nchan + 2)
    cind = np.floor(imel(cbin) / fs * nfft).astype(int) + 1
    mfb = np.zeros((len(fbin), nchan))
    for i in range(nchan):
        a, b, c = cind[i], cind[i + 1], cind[i + 2]
        mfb[a:b, i] = (cbin[i] - fbin[a:b]) / (cbin[i] - cbin[i + 1])
        mfb[b:c, i] = (cbin[i + 2] - fbin[b:c]) / (cbin[i + 2] - cbin[i + 1])
    return mfb


def dither_stream(n, level=8):
    """The noise add_dither draws after ``np.random.seed(3)`` (vbx_segmenter.py:84-85):
    a prefix of ONE fixed legacy-MT19937 sequence, independent of the signal."""
    np.random.seed(3)
    return level * (np.random.rand(n) * 2 - 1)


def quantise(signal):
    """(signal * 2**15).astype(int) -- truncation toward zero (vbx_segmenter.py:85)."""
    return (np.asarray(signal, dtype=np.float64) * 2 ** 15).astype(int)


def raw_fbank(signal):
    """Everything of get_features before CMVN: [M, 64] float64."""
    x = quantise(signal) + dither_stream(len(signal))
    seg = np.r_[x[NOVERLAP // 2 - 1::-1], x, x[-1:-FRAME // 2 - 1:-1]]        # mirror pad 120 / 200
    M = (len(seg) - FRAME) // (FRAME - NOVERLAP) + 1
    fr = np.lib.stride_tricks.as_strided(seg, shape=(M, FRAME), strides=(seg.strides[0] * (FRAME - NOVERLAP), seg.strides[0])).copy()
    fr -= fr.mean(axis=1)[:, np.newaxis]                                      # ZMEANSOURCE
    fr = fr - np.c_[fr[..., :1], fr[..., :-1]] * 0.97                         # per-frame pre-emphasis
    fr *= povey_window()
    sp = np.fft.rfft(fr, NFFT)
    p = sp.real ** 2 + sp.imag ** 2                                           # USEPOWER=True -> power
    return np.log(np.maximum(1.0, np.dot(p, mel_bank())))


def cmvn_floating(x, LC=150, RC=149):
    """cmvn_floating_kaldi(x, 150, 149, norm_vars=False)."""
    N, dim = x.shape
    win_len = min(len(x), LC + RC + 1)
    win_start = np.maximum(np.minimum(np.arange(-LC, N - LC), N - win_len), 0)
    f = np.r_[np.zeros((1, dim)), np.cumsum(x, 0)]
    return x - (f[win_start + win_len] - f[win_start]) / win_len


def get_features(signal, LC=150, RC=149):
    return cmvn_floating(raw_fbank(signal), LC, RC).astype(np.float32)


def window_plan(M):
    """(start, length, is_tail) of every window VBxExtractor.__call__ embeds
    (vbx_segmenter.py:222-243): regular 144-frame windows every 24 frames while
    start < M - 144, then one tail window fea[last_start + 24:] if >= 10 frames remain."""
    plan = []
    start = 0
    for start in range(0, M - WINLEN, STEP):
        plan.append((start, WINLEN, False))
    if M - start - STEP >= 10:
        plan.append((start + STEP, M - (start + STEP), True))
    return plan


# ------------------------------------------------------------------ ResNet101 (resnet.py)
def synthetic_resnet101_state(seed=0, feat_dim=FEAT_DIM, embed_dim=EMBED_DIM, m=32):
    """Seeded state_dict with the parameter names / shapes of resnet.ResNet101(feat_dim, embed_dim)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin, k):
        sd[name + '.weight'] = torch.randn(cout, cin, k, k, generator=g) * float(np.sqrt(2.0 / (cin * k * k)))

    def bn(name, c):
        sd[name + '.weight'] = torch.rand(c, generator=g) * 0.4 + 0.8
        sd[name + '.bias'] = torch.randn(c, generator=g) * 0.1
        sd[name + '.running_mean'] = torch.randn(c, generator=g) * 0.1
        sd[name + '.running_var'] = torch.rand(c, generator=g) + 0.5

    conv('conv1', m, 1, 3); bn('bn1', m)
    inp = m
    for li, (planes, nb, stride) in enumerate(zip((m, 2 * m, 4 * m, 8 * m), (3, 4, 23, 3), (1, 2, 2, 2)), start=1):
        for b in range(nb):
            p = 'layer%d.%d' % (li, b)
            s = stride if b == 0 else 1
            conv(p + '.conv1', planes, inp, 1); bn(p + '.bn1', planes)
            conv(p + '.conv2', planes, planes, 3); bn(p + '.bn2', planes)
            conv(p + '.conv3', 4 * planes, planes, 1); bn(p + '.bn3', 4 * planes)
            # keep residual streams well scaled through 33 blocks
            sd[p + '.bn3.weight'] *= 0.3
            if s != 1 or inp != 4 * planes:
                conv(p + '.shortcut.0', 4 * planes, inp, 1); bn(p + '.shortcut.1', 4 * planes)
            inp = 4 * planes
    d = (feat_dim // 8) * m * 16 * 4
    sd['embedding.weight'] = torch.randn(embed_dim, d, generator=g) * float(np.sqrt(1.0 / d))
    sd['embedding.bias'] = torch.randn(embed_dim, generator=g) * 0.05
    return sd


class ResNet101Oracle:
    def __init__(self, state, threads=None):
        self.sd = {k: v.float() for k, v in state.items()}
        self.threads = threads

    def _bn(self, x, p):
        sd = self.sd
        return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'], sd[p + '.bias'],
                            training=False, eps=1e-5)

    @torch.no_grad()
    def forward(self, x):
        """x: [n, 64, T] float32 (resnet.py:115-130) -> [n, 256]."""
        if self.threads:
            torch.set_num_threads(self.threads)
        sd = self.sd
        out = F.relu(self._bn(F.conv2d(x.unsqueeze(1), sd['conv1.weight'], padding=1), 'bn1'))
        for li, (nb, stride) in enumerate(zip((3, 4, 23, 3), (1, 2, 2, 2)), start=1):
            for b in range(nb):
                p = 'layer%d.%d' % (li, b)
                s = stride if b == 0 else 1
                y = F.relu(self._bn(F.conv2d(out, sd[p + '.conv1.weight']), p + '.bn1'))
                y = F.relu(self._bn(F.conv2d(y, sd[p + '.conv2.weight'], stride=s, padding=1), p + '.bn2'))
                y = self._bn(F.conv2d(y, sd[p + '.conv3.weight']), p + '.bn3')
                if (p + '.shortcut.0.weight') in sd:
                    sc = self._bn(F.conv2d(out, sd[p + '.shortcut.0.weight'], stride=s), p + '.shortcut.1')
                else:
                    sc = out
                out = F.relu(y + sc)
        mean = torch.mean(out, dim=-1)
        meansq = torch.mean(out * out, dim=-1)
        std = torch.sqrt(meansq - mean ** 2 + 1e-10)
        feat = torch.cat((torch.flatten(mean, start_dim=1), torch.flatten(std, start_dim=1)), 1)
        return feat @ sd['embedding.weight'].t() + sd['embedding.bias']

    def get_embedding(self, fea):
        """OnnxBackendExtractor.get_embedding semantics (vbx_segmenter.py:262-266): fea [T, 64] -> [256]."""
        x = torch.from_numpy(np.ascontiguousarray(fea, dtype=np.float32).T[np.newaxis])
        return self.forward(x)[0].numpy()


def extract_xvectors(fea, net, basename, duration):
    """VBxExtractor.__call__ (vbx_segmenter.py:217-246) with `net.get_embedding`."""
    out = []
    M = len(fea)
    for start, length, is_tail in window_plan(M):
        x = net.get_embedding(fea[start:start + length])
        if np.isnan(x).any():
            continue
        if not is_tail:
            key = '%s_%08d-%08d' % (basename, start, start + WINLEN)
            seg = (round(start / 100.0, 3), round(start / 100.0 + WINLEN / 100.0, 3))
        else:
            key = '%s_%08d-%08d' % (basename, start, M)
            seg = (round(start / 100.0, 3), round(duration, 3))
        out.append((key, seg, x * 10))
    return out
