#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REAL reference modules.

Run in the build container only (needs /root/reference, read-only).  The
reference package cannot be imported (``__init__`` pulls TensorFlow), so the
numpy-only modules are loaded by file path.  For every fixture the script
(1) asserts the oracle restatement reproduces the reference output
bit-for-bit, then (2) stores a compact golden so the same check runs on the
GPU box where /root/reference does not exist.

    python tests/golden/make_golden.py
"""
import hashlib
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference/inaSpeechSegmenter'


def load_ref(name):
    spec = importlib.util.spec_from_file_location('ref_' + name, os.path.join(REF, name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from oracle import sidekit_oracle as sk
    from oracle import segmenter_oracle as so
    from oracle import viterbi_oracle as vo
    ref_mfcc = load_ref('sidekit_mfcc')
    ref_vit = load_ref('pyannote_viterbi')
    ref_vu = load_ref('viterbi_utils')
    out = {}

    # ---- filterbank / window tables ---------------------------------------
    fb_ref = ref_mfcc.trfbank(16000, 512, 100, 8000, 0, 24)[0]
    fb = sk.mel_filterbank()[0]
    assert fb.dtype == np.float32 and np.array_equal(fb, fb_ref), 'trfbank restatement differs'
    out['fbank'] = fb_ref
    print('trfbank: identical, nnz=%d' % np.count_nonzero(fb))

    # ---- front-end on fixtures + synthetic --------------------------------
    rng = np.random.default_rng(20260922)
    synth = (rng.standard_normal(16000 * 7) * 0.1).astype(np.float32)
    synth[16000:2 * 16000] = 0.0                       # an exact-silence span -> -inf
    t = np.arange(3 * 16000) / 16000.
    synth[3 * 16000:6 * 16000] += (0.3 * np.sin(2 * np.pi * 220 * t) * (1 + 0.5 * np.sin(2 * np.pi * 4 * t))).astype(np.float32)
    synth = (np.round(synth * 32768).clip(-32768, 32767) / 32768).astype(np.float32)
    out['synth_sig'] = synth
    signals = {
        'musanmix': so.read_wav_16k_mono(os.path.join(HERE, 'media', 'musanmix.wav')),
        'silence2sec': so.read_wav_16k_mono(os.path.join(HERE, 'media', 'silence2sec.wav')),
        'lamartine': so.read_wav_16k_mono(os.path.join(HERE, 'media', 'lamartine.wav')),
        'synth': synth,
    }
    for name, sig in signals.items():
        with np.errstate(divide='ignore'):
            _, loge_r, _, mspec_r = ref_mfcc.mfcc(sig.astype(np.float32), get_mspec=True)
        mspec_o, loge_o = sk.logmel_loge(sig)
        assert loge_r.dtype == np.float32 and mspec_r.dtype == np.float32
        assert np.array_equal(loge_r, loge_o, equal_nan=True), name
        assert np.array_equal(mspec_r, mspec_o, equal_nan=True), name
        L = len(loge_r)
        assert L == sk.num_frames(len(sig))
        # compact golden: head/tail rows verbatim + digest of the full arrays
        keep = np.r_[0:min(L, 300), max(L - 100, 0):L]
        out[name + '_nsamp'] = np.int64(len(sig))
        out[name + '_rows'] = keep
        out[name + '_mspec'] = mspec_r[keep]
        out[name + '_loge'] = loge_r[keep]
        out[name + '_sha_mspec'] = np.array(digest(mspec_r))
        out[name + '_sha_loge'] = np.array(digest(loge_r))
        print('%-12s L=%d  mspec/loge identical to reference' % (name, L))
        # energy activity (oracle glue) -> stored for the product tests
        segs = so.energy_segments(loge_r)
        out[name + '_energy_segs'] = np.array([(1 if lab == 'energy' else 0, a, b) for lab, a, b in segs], dtype=np.int64)

    # ---- viterbi -----------------------------------------------------------
    cases = []
    for k, T, seed in [(2, 1, 1), (2, 2, 2), (2, 997, 3), (3, 1500, 4), (3, 1, 5), (2, 5000, 6), (3, 4001, 7)]:
        r = np.random.default_rng(seed)
        if k == 2 and seed % 2 == 0:
            raw = r.random(T) > 0.5
            raw[T // 3: T // 3 + 40] = True
            em = ref_vu.pred2logemission(raw)
            tr = ref_vu.log_trans_exp(150, cost0=-5)
            assert np.array_equal(em, vo.pred2logemission(raw)) and np.array_equal(tr, vo.log_trans_exp(150, cost0=-5))
        else:
            p = r.random((T, k)).astype(np.float32) ** 4 + 1e-6
            p /= p.sum(1, keepdims=True)
            # piecewise-constant dominant class so smoothing matters
            dom = np.repeat(r.integers(0, k, T // 50 + 1), 50)[:T]
            p[np.arange(T), dom] += 1.0
            p = (p / p.sum(1, keepdims=True)).astype(np.float32)
            p[T // 2] = 0.5                                 # the non-finite-patch override
            em = np.log(p)
            tr = ref_vu.diag_trans_exp(80, k)
            assert np.array_equal(tr, vo.diag_trans_exp(80, k))
        st_ref = ref_vit.viterbi_decoding(em.copy(), tr)
        assert np.array_equal(st_ref, vo.viterbi_numpy(em, tr)), (k, T, seed)
        assert np.array_equal(st_ref, vo.viterbi_c(em, tr)), (k, T, seed)
        i = len(cases)
        out['vit%d_em' % i], out['vit%d_tr' % i], out['vit%d_st' % i] = em, tr, st_ref.astype(np.int8)
        cases.append((k, T))
    out['vit_ncases'] = np.int64(len(cases))
    print('viterbi: %d cases identical (numpy + C restatements)' % len(cases))

    # ---- VBx front-end -----------------------------------------------------
    try:
        from oracle import vbx_oracle as vx
    except ImportError:
        vx = None
    if vx is not None:
        ref_fv = load_ref('features_vbx')

        def ref_get_features(signal, LC=150, RC=149):
            # vbx_segmenter.py:72-89 cannot be imported (onnxruntime/keras); these
            # eleven lines ARE the reference recipe, executed with its own helpers.
            window = ref_fv.povey_window(400)
            fbank_mx = ref_fv.mel_fbank_mx(400, 16000, NUMCHANS=64, LOFREQ=20.0, HIFREQ=7600, htk_bug=False)
            np.random.seed(3)
            signal = ref_fv.add_dither((signal * 2 ** 15).astype(int))
            seg = np.r_[signal[240 // 2 - 1::-1], signal, signal[-1:-400 // 2 - 1:-1]]
            fea = ref_fv.fbank_htk(seg, window, 240, fbank_mx, USEPOWER=True, ZMEANSOURCE=True)
            return ref_fv.cmvn_floating_kaldi(fea, LC, RC, norm_vars=False).astype(np.float32)

        sig64 = so.read_wav_16k_mono(os.path.join(HERE, 'media', 'lamartine.wav'), dtype='float64')
        for name, s in (('lamartine', sig64), ('synth', synth.astype(np.float64)), ('short', synth[:16000 * 2 + 77].astype(np.float64))):
            fr = ref_get_features(s)
            fo = vx.get_features(s)
            assert fr.dtype == np.float32 and np.array_equal(fr, fo), name
            M = len(fr)
            keep = np.r_[0:min(M, 200), max(M - 60, 0):M]
            out['vbx_%s_rows' % name] = keep
            out['vbx_%s_fea' % name] = fr[keep]
            out['vbx_%s_sha' % name] = np.array(digest(fr))
            out['vbx_%s_M' % name] = np.int64(M)
            print('vbx %-10s M=%d identical to reference' % (name, M))

    # ---- ResNet101 architecture restatement vs the real resnet.py --------------
    if vx is not None:
        import torch
        ref_resnet = load_ref('resnet')
        sd = vx.synthetic_resnet101_state(seed=5)
        net = ref_resnet.ResNet101(feat_dim=64, embed_dim=256)
        missing, unexpected = net.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.endswith('num_batches_tracked') for k in missing), (missing, unexpected)
        net.eval()
        g = torch.Generator().manual_seed(9)
        x = torch.randn(2, 64, 144, generator=g)
        xs = torch.randn(1, 64, 131, generator=g)                 # a tail-window length
        with torch.no_grad():
            y_ref, ys_ref = net(x.clone()), net(xs.clone())
        orc = vx.ResNet101Oracle(sd)
        y, ys = orc.forward(x), orc.forward(xs)
        assert torch.equal(y, y_ref) and torch.equal(ys, ys_ref), (float((y - y_ref).abs().max()))
        out['resnet_x'], out['resnet_y'] = x.numpy(), y_ref.numpy()
        out['resnet_xs'], out['resnet_ys'] = xs.numpy(), ys_ref.numpy()
        print('resnet101: functional restatement identical to resnet.py (|y| mean %.3f)' % float(y_ref.abs().mean()))
        # window plan vs the reference loop (vbx_segmenter.py:222-243 re-typed as a generator of (start, stop))
        for M in (9, 10, 33, 34, 143, 144, 145, 167, 168, 169, 200, 1464, 1465):
            ref_plan, start = [], 0
            for start in range(0, M - 144, 24):
                ref_plan.append((start, start + 144))
            if M - start - 24 >= 10:
                ref_plan.append((start + 24, M))
            assert [(a, a + n) for a, n, _ in vx.window_plan(M)] == ref_plan, M

    np.savez_compressed(os.path.join(HERE, 'reference_golden.npz'), **out)
    print('wrote', os.path.join(HERE, 'reference_golden.npz'))


if __name__ == '__main__':
    main()
