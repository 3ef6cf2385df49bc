"""VBx x-vector path on B200: host mirror of the numeric part of the reference's
``vbx_segmenter.py`` (constants :21-25, ``get_features`` :72-89, the
``VBxExtractor`` plugin ABC :205-246 and a backend :249-266).

``B200BackendExtractor`` is a drop-in for ``OnnxBackendExtractor``: same
``get_embedding(fea[T,64]) -> [256]`` and ``__call__(basename, fea, duration)``
contract, but ``__call__`` embeds all windows in batches on the GPU instead of
one ONNX call per 0.24 s of audio.  ``VoiceFemininityScoring`` (:92-202, SURVEY
N3) is mirrored at the end of the file: the MLP runs on device (``iss_mlp_*``),
the pyannote interval bookkeeping is restated in plain Python.
"""
import ctypes
import logging
import os
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import _lib
from .engine import Context, _stream_ptr

logger = logging.getLogger(__name__)

STEP = 24
WINLEN = 144
FEAT_DIM = 64
EMBED_DIM = 256
SR = 16000
NUM_BLOCKS = (3, 4, 23, 3)          # resnet.py:133-135 (ResNet101)
M_CHANNELS = 32                     # resnet.py:79


# ------------------------------------------------------------------ host tables (features_vbx.py)
def povey_window(winlen=400):
    """features_vbx.py:123-124."""
    return np.power(0.5 - 0.5 * np.cos(np.linspace(0, 2 * np.pi, winlen)), 0.85)


def mel_fbank_htk64(nfft=512, fs=SR, nchan=FEAT_DIM, lofreq=20.0, hifreq=7600.0):
    """mel_fbank_mx(400, 16000, NUMCHANS=64, LOFREQ=20, HIFREQ=7600, htk_bug=False)
    (features_vbx.py:31-59): float64 [257, 64]."""
    warp = lambda x: 1127. * np.log(1. + x / 700.)         # noqa: E731
    unwarp = lambda x: (np.exp(x / 1127.) - 1.) * 700.     # noqa: E731
    bins = warp(np.arange(nfft / 2 + 1, dtype=float) * fs / nfft)
    cent = np.linspace(warp(lofreq), warp(hifreq), nchan + 2)
    idx = np.floor(unwarp(cent) / fs * nfft).astype(int) + 1
    mx = np.zeros((len(bins), nchan))
    for i in range(nchan):
        mx[idx[i]:idx[i + 1], i] = (cent[i] - bins[idx[i]:idx[i + 1]]) / (cent[i] - cent[i + 1])
        mx[idx[i + 1]:idx[i + 2], i] = (cent[i + 2] - bins[idx[i + 1]:idx[i + 2]]) / (cent[i + 2] - cent[i + 1])
    return mx


class DitherCache:
    """Device-resident prefix of the dither the reference draws with
    ``np.random.seed(3); 8 * (np.random.rand(n) * 2 - 1)`` (vbx_segmenter.py:84-85,
    features_vbx.py:127-128).  The stream does not depend on the signal, so it is
    generated once (legacy MT19937 via RandomState(3): same doubles, no global
    RNG side effect) and kept in HBM; longer requests regenerate a longer prefix."""

    def __init__(self, device, level=8):
        self.device, self.level, self.buf = device, level, None

    def get(self, n):
        if self.buf is None or self.buf.numel() < n:
            cap = max(n, 16000 * 60)
            host = self.level * (np.random.RandomState(3).rand(cap) * 2 - 1)
            self.buf = torch.from_numpy(host).to(self.device)
        return self.buf


class VbxFrontEnd:
    def __init__(self, ctx):
        self.ctx = ctx
        self._fbank = np.ascontiguousarray(mel_fbank_htk64(), dtype=np.float64)
        self._window = np.ascontiguousarray(povey_window(400), dtype=np.float64)
        _lib.check(_lib.load().iss_vbx_upload_tables(ctx.handle, _lib.ptr(self._fbank), _lib.ptr(self._window)),
                   'iss_vbx_upload_tables')
        self.dither = DitherCache(ctx.device)

    def __call__(self, pcm, dither=True, stream=None):
        """pcm: CUDA int16 or float32 [n] -> CUDA float32 [M, 64] (CMVN'd log-mel)."""
        assert pcm.is_cuda and pcm.dim() == 1 and pcm.is_contiguous()
        fmt = {torch.int16: _lib.PCM_S16, torch.float32: _lib.PCM_F32}[pcm.dtype]
        lib = _lib.load()
        n = pcm.numel()
        M = lib.iss_vbx_num_frames(n)
        fea = torch.empty((M, FEAT_DIM), dtype=torch.float32, device=pcm.device)
        if M == 0:
            return fea
        work = self.ctx.workspace('vbx', lib.iss_vbx_work_bytes(n))
        d = self.dither.get(n) if dither else None
        _lib.check(lib.iss_vbx_features(self.ctx.handle, _lib.ptr(pcm), fmt, n, _lib.ptr(d), _lib.ptr(fea),
                                        _lib.ptr(work), _stream_ptr(pcm.device, stream)), 'iss_vbx_features')
        return fea


_frontends = {}


def get_features(signal, LC=150, RC=149, device=0):
    """Same contract as the reference (vbx_segmenter.py:72-89): float64 (or
    float32 / int16) signal in [-1, 1) -> float32 [M, 64] numpy array."""
    assert (LC, RC) == (150, 149), 'only the reference window (150, 149) is built into the kernel'
    if device not in _frontends:
        _frontends[device] = VbxFrontEnd(Context(device))
    fe = _frontends[device]
    sig = np.asarray(signal)
    if sig.dtype == np.int16:
        pcm = torch.from_numpy(np.ascontiguousarray(sig))
    else:
        f32 = sig.astype(np.float32)
        if not np.array_equal(f32.astype(np.float64), sig.astype(np.float64)):
            raise ValueError('signal is not exactly representable in float32 (the decoder yields PCM16/float32 data)')
        pcm = torch.from_numpy(np.ascontiguousarray(f32))
    return fe(pcm.to(fe.ctx.device)).cpu().numpy()


# ------------------------------------------------------------------ ResNet101 weights -> blob
def resnet_blob_from_state(sd, m=M_CHANNELS, num_blocks=NUM_BLOCKS, eps=1e-5):
    """PyTorch state_dict of resnet.ResNet101 (``raw_81.pth`` layout, resnet.py:78-113)
    -> the float32 blob iss_resnet_create expects (BatchNorm folded to scale/shift)."""
    g = lambda k: np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], 'detach') else sd[k], dtype=np.float32)   # noqa: E731
    parts = []

    def conv_bn(conv, bn):
        w = g(conv + '.weight')                                   # [cout, cin, kh, kw]
        parts.append(np.ascontiguousarray(w.transpose(2, 3, 1, 0)).ravel())       # [kh][kw][cin][cout]
        scale = g(bn + '.weight') / np.sqrt(g(bn + '.running_var') + np.float32(eps))
        shift = g(bn + '.bias') - g(bn + '.running_mean') * scale
        parts.append(scale.astype(np.float32))
        parts.append(shift.astype(np.float32))

    conv_bn('conv1', 'bn1')
    for li, nb in enumerate(num_blocks, start=1):
        for b in range(nb):
            p = 'layer%d.%d' % (li, b)
            conv_bn(p + '.conv1', p + '.bn1')
            conv_bn(p + '.conv2', p + '.bn2')
            conv_bn(p + '.conv3', p + '.bn3')
            if (p + '.shortcut.0.weight') in sd:
                conv_bn(p + '.shortcut.0', p + '.shortcut.1')
    parts.append(np.ascontiguousarray(g('embedding.weight').T).ravel())           # [in][embed]
    parts.append(g('embedding.bias'))
    return np.concatenate(parts).astype(np.float32)


# ------------------------------------------------------------------ the plugin ABC (vbx_segmenter.py:205-246)
class VBxExtractor(ABC):
    """VBxExtractor is an abstract class performing xvector extraction."""

    @abstractmethod
    def __init__(self):
        pass

    def __call__(self, basename, fea, duration):
        """One get_embedding() call per planned window (vbx_segmenter.py:217-246): NaN embeddings are dropped with a
        warning, the tail window ends at `duration`, embeddings are returned x10."""
        out = []
        for start, length, is_tail in window_plan(len(fea)):
            stop = start + length
            key = '%s_%08d-%08d' % (basename, start, stop)
            x = self.get_embedding(fea[start:stop])
            if np.isnan(x).any():
                logger.warning('NaN found, not processing: %s%s' % (key, os.linesep))
                continue
            t0 = round(start / 100.0, 3)
            t1 = round(duration, 3) if is_tail else round(start / 100.0 + WINLEN / 100.0, 3)
            out.append((key, (t0, t1), x * 10))
        return out


def window_plan(M):
    """(start, length, is_tail) for every window of VBxExtractor.__call__ (vbx_segmenter.py:222-243)."""
    plan, start = [], 0
    for start in range(0, M - WINLEN, STEP):
        plan.append((start, WINLEN, False))
    if M - start - STEP >= 10:
        plan.append((start + STEP, M - (start + STEP), True))
    return plan


class B200BackendExtractor(VBxExtractor):
    """ResNet101 x-vector extractor on libiss_b200 (K5)."""

    def __init__(self, state_dict=None, device=0, ctx=None, onnx_path=None):
        """Weights come from (first match): ``state_dict`` (resnet.py layout), ``onnx_path``, the reference's
        production asset ``final.onnx`` (vbx_segmenter.py:249-266) or ``raw_81.pth`` (:271-288) in the model
        directories (models.find_model_file).  The ONNX file is read by onnx_reader (no onnx runtime)."""
        self.ctx = ctx if ctx is not None else Context(device)
        m, feat_dim, embed_dim, num_blocks = M_CHANNELS, FEAT_DIM, EMBED_DIM, NUM_BLOCKS
        if state_dict is None:
            from .models import find_model_file
            onnx_path = onnx_path or find_model_file('final.onnx')
            if onnx_path is not None:
                from .onnx_reader import resnet_blob_from_onnx
                blob, m, feat_dim, embed_dim, num_blocks = resnet_blob_from_onnx(onnx_path)
                if (m, feat_dim, embed_dim, tuple(num_blocks)) != (M_CHANNELS, FEAT_DIM, EMBED_DIM, NUM_BLOCKS):
                    raise ValueError('%s is not the VBx ResNet101 (m=%d, feat=%d, embed=%d, blocks=%r)'
                                     % (onnx_path, m, feat_dim, embed_dim, num_blocks))
            else:
                path = find_model_file('raw_81.pth')
                if path is None:
                    raise FileNotFoundError('neither final.onnx nor raw_81.pth found (the reference fetches them from its '
                                            'GitHub release, remote_utils.py:5,13-14; this build does no network access)')
                state_dict = torch.load(path, map_location='cpu')
                state_dict = state_dict.get('state_dict', state_dict)
        if state_dict is not None:
            blob = resnet_blob_from_state(state_dict)
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        nb = (ctypes.c_int * 4)(*NUM_BLOCKS)
        lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(lib.iss_resnet_create(self.ctx.handle, _lib.ptr(blob), blob.size, M_CHANNELS, FEAT_DIM, EMBED_DIM,
                                         ctypes.cast(nb, ctypes.c_void_p), ctypes.byref(h)), 'iss_resnet_create')
        self.handle = h
        self.flops_per_window = lib.iss_resnet_flops_per_window(h, WINLEN)

    def embed_windows(self, fea, starts, win_len, stream=None):
        """fea: CUDA float32 [M, 64]; starts: window start rows -> CUDA float32 [n, 256]."""
        assert fea.is_cuda and fea.dtype == torch.float32 and fea.is_contiguous()
        starts = np.ascontiguousarray(starts, dtype=np.int32)
        n = len(starts)
        emb = torch.empty((n, EMBED_DIM), dtype=torch.float32, device=fea.device)
        if n == 0:
            return emb
        lib = _lib.load()
        work = self.ctx.workspace('resnet', lib.iss_resnet_workspace_bytes(self.handle, n, int(win_len)))
        _lib.check(lib.iss_resnet_embed(self.ctx.handle, self.handle, _lib.ptr(fea), fea.shape[0], _lib.ptr(starts), n,
                                        int(win_len), _lib.ptr(emb), _lib.ptr(work), work.numel(),
                                        _stream_ptr(fea.device, stream)), 'iss_resnet_embed')
        return emb

    def get_embedding(self, fea):
        """OnnxBackendExtractor.get_embedding contract (vbx_segmenter.py:262-266): fea [T, 64] -> [256]."""
        t = torch.from_numpy(np.ascontiguousarray(fea, dtype=np.float32)).to(self.ctx.device)
        return self.embed_windows(t, [0], t.shape[0]).cpu().numpy()[0]

    def __call__(self, basename, fea, duration):
        """Same output as VBxExtractor.__call__, windows embedded in GPU batches."""
        if isinstance(fea, np.ndarray):
            fea = torch.from_numpy(np.ascontiguousarray(fea, dtype=np.float32)).to(self.ctx.device)
        M = fea.shape[0]
        plan = window_plan(M)
        reg = [s for s, n, tail in plan if not tail]
        emb = self.embed_windows(fea, reg, WINLEN).cpu().numpy() if reg else np.zeros((0, EMBED_DIM), np.float32)
        out = []
        for i, start in enumerate(reg):
            key = f'{basename}_{start:08}-{(start + WINLEN):08}'
            if np.isnan(emb[i]).any():
                logger.warning(f'NaN found, not processing: {key}{os.linesep}')
                continue
            out.append((key, (round(start / 100.0, 3), round(start / 100.0 + WINLEN / 100.0, 3)), emb[i]))
        for start, n, tail in plan:
            if tail:
                x = self.embed_windows(fea, [start], n).cpu().numpy()[0]
                key = f'{basename}_{start:08}-{M:08}'
                if np.isnan(x).any():
                    logger.warning(f'NaN found, not processing: {key}{os.linesep}')
                else:
                    out.append((key, (round(start / 100.0, 3), round(duration, 3)), x))
        return [(key, seg, x * 10) for key, seg, x in out]


# ------------------------------------------------------------------ VoiceFemininityScoring (vbx_segmenter.py:28-202)
def _overlap(start, stop, speech):
    """Total duration of [start, stop] covered by the (disjoint) speech intervals: what
    ``Timeline([Segment(start, stop)]).crop(vad_timeline).duration()`` yields (:140)."""
    return sum(max(0.0, min(stop, e) - max(start, b)) for b, e in speech)


def is_mid_speech(start, stop, speech):
    """True if the window's midpoint lies strictly inside a speech segment (:28-37)."""
    m = (start + stop) / 2
    return any(b < m < e for b, e in speech)


def add_needed_vectors(xvectors, t_mid):
    """Keep at least 50 % of the windows whose midpoint is in speech (:40-52): when the
    overlap threshold removed too many, the windows are taken back in decreasing overlap order.
    (The reference sorts a ragged object array and reads ``Segment.stop``, which pyannote's Segment
    does not define; this implements the documented intent with the same selection rule: sort by
    overlap ratio descending, skip the first len(xvectors) entries, take the next ``diff``.)"""
    min_pred = round(0.5 * len(t_mid))
    if len(xvectors) < min_pred:
        order = np.argsort(np.array([t[0] for t in t_mid], dtype=np.float64))[::-1]
        ranked = [t_mid[i] for i in order]
        diff = min_pred - len(xvectors)
        for _, k, (s0, s1), x in ranked[len(xvectors):len(xvectors) + diff]:
            xvectors.append((k, (s0, s1), x))
    return xvectors


def get_femininity_score(g_preds):
    """Share of retained windows predicted feminine (p >= 0.5) (:55-61).  The reference goes through a
    pyannote Annotation keyed by segment, so a later window with identical bounds replaces an earlier one."""
    by_segment = {}
    for start, stop, p in g_preds:
        by_segment[(start, stop)] = bool(np.all(np.asarray(p) >= 0.5))
    return sum(by_segment.values()) / len(by_segment)


class VoiceFemininityScoring:
    """Same contract as the reference class (vbx_segmenter.py:92-202): ``__call__(fpath) ->
    (score, speech_duration, nb_vectors)``.  VAD = the B200 ``Segmenter('smn', detect_gender=False)``,
    x-vectors = K4 + K5, the gender MLP runs on device through ``iss_mlp_forward``; the interval
    bookkeeping the reference delegates to pyannote.core is plain Python here.
    Extra optional arguments: ``device``, ``ffmpeg`` and ``models`` = {'vad': (config, weights),
    'mlp': (config, weights), 'resnet': state_dict} to bypass the model-file lookup."""

    def __init__(self, gd_model_criteria="bgc", backend='onnx', device=0, ffmpeg='ffmpeg', models=None):
        assert backend in ['onnx', 'b200'], "Backend should be 'onnx' (served by the B200 extractor) or 'b200'."
        from .engine import MlpModel
        from .models import find_model_file, load_model_file
        from .segmenter import Segmenter
        models = models or {}
        assert gd_model_criteria in ["bgc", "vfp"], "Gender detection model Criteria must be 'bgc' (default) or 'vfp'"
        if gd_model_criteria == "bgc":
            gd_model, self.vad_thresh = "interspeech2023_all.hdf5", 0.7
        else:
            gd_model, self.vad_thresh = "interspeech2023_cvfr.hdf5", 0.62
        self.vad = Segmenter(vad_engine='smn', detect_gender=False, ffmpeg=ffmpeg, device=device,
                             models={'vad': models['vad']} if 'vad' in models else None)
        self.ctx = self.vad.ctx
        self.xvector_model = B200BackendExtractor(state_dict=models.get('resnet'), ctx=self.ctx)
        mlp = models.get('mlp')
        if mlp is None:
            path = find_model_file(gd_model)
            if path is None:
                raise FileNotFoundError('%s not found (reference asset, remote_utils.py:4-15)' % gd_model)
            mlp = load_model_file(path)
        self.gender_detection_mlp_model = MlpModel(self.ctx, mlp[0], mlp[1], EMBED_DIM)
        self.frontend = VbxFrontEnd(self.ctx)
        self.ffmpeg = ffmpeg

    def apply_vad(self, xvectors, speech):
        """vbx_segmenter.py:129-145 with `speech` = [(start, end)] of the 'speech' VAD segments."""
        midpoint_seg, kept = [], []
        for key, (start, stop), x in xvectors:
            if is_mid_speech(start, stop, speech):
                ratio = _overlap(start, stop, speech) / (stop - start)
                if ratio >= self.vad_thresh:
                    kept.append((key, (start, stop), x))
                midpoint_seg.append((ratio, key, (start, stop), x))
        return add_needed_vectors(kept, midpoint_seg)

    def score_signal(self, sig, basename='signal'):
        """The body of __call__ on a decoded 16 kHz mono signal (numpy int16 / float)."""
        sig = np.asarray(sig)
        duration = len(sig) / SR
        vad_seg = self.vad.segment_signal(sig if sig.dtype in (np.int16, np.float32) else sig.astype(np.float32))
        speech = [(b, e) for lab, b, e in vad_seg if lab == 'speech']
        speech_duration = sum(e - b for b, e in speech)
        if not speech_duration:
            return None, speech_duration, 0
        pcm = torch.from_numpy(np.ascontiguousarray(sig if sig.dtype == np.int16 else sig.astype(np.float32))).to(self.ctx.device)
        features = self.frontend(pcm)
        x_vectors = self.xvector_model(basename, features, duration)
        x_vectors = self.apply_vad(x_vectors, speech)
        x = np.asarray([x for _, _, x in x_vectors])
        gender_pred = self.gender_detection_mlp_model.predict(x, verbose=0)
        if len(gender_pred) > 1:
            gender_pred = np.squeeze(gender_pred)
        g = [(seg[0], seg[1], p) for (_, seg, _), p in zip(x_vectors, gender_pred)]
        return get_femininity_score(g), speech_duration, len(g)

    def __call__(self, fpath):
        """Voice femininity score of a media file (vbx_segmenter.py:147-202)."""
        from .io import media2sig16kmono
        basename = os.path.splitext(os.path.basename(fpath))[0]
        sig = media2sig16kmono(fpath, ffmpeg=self.ffmpeg, dtype='float32', return_int16=True)
        return self.score_signal(sig, basename)
