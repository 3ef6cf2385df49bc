"""CPU models of the index algebra inside the K1 feature kernel (csrc/feat_sidekit.cu, second pass of round 2):

* the even/odd split of the 512-point real FFT takes Z[256 - k] from a partner lane with register shuffles:
  partner lane / register selection, the lane-ordered twiddle table and the conflict-free layout of the power
  spectrum are checked against numpy's rfft on a 32-lane model;
* the balanced mel tasks (host code of iss_sidekit_upload_tables) cover every non-zero of every filter exactly once
  with at most 32 tasks.
"""
import numpy as np

from inaspeechsegmenter_b200 import sidekit_mfcc as sm


def _bitrev2(r):
    return ((r & 1) << 1) | ((r >> 1) & 1)


def _k(lane, a):
    """bin held by `lane` in register `a` after warp_fft256_reg (fft256r.cuh)"""
    return (lane >> 2) + 8 * a + 64 * _bitrev2(lane & 3)


def test_split_by_shuffles_reproduces_rfft():
    rng = np.random.default_rng(0)
    v = rng.standard_normal(512)
    Z = np.fft.fft(v[0::2] + 1j * v[1::2])
    X = np.fft.rfft(v)
    reg = np.array([[Z[_k(L, a)] for a in range(8)] for L in range(32)])
    assert sorted(_k(L, a) for L in range(32) for a in range(8)) == list(range(256))
    P = np.full(257, np.nan)
    for a in range(8):
        send, src = np.zeros(32, complex), np.zeros(32, int)
        for L in range(32):
            k1, r = L >> 2, L & 3
            cls0 = k1 == 0
            send[L] = reg[L, (8 - a) & 7] if cls0 else reg[L, 7 - a]            # the SOURCE lane picks the register
            gen = (3 - r) if cls0 else ((((8 - k1) & 7) << 2) | (3 - r))
            src[L] = ((r ^ (r >> 1)) if cls0 else gen) if a == 0 else gen       # {0, 1, 3, 2}[r]
        for L in range(32):
            k = _k(L, a)
            partner = send[src[L]]
            assert np.isclose(partner, Z[(256 - k) % 256]), (L, a, k)
            me, c = reg[L, a], np.conj(partner)
            x = 0.5 * (me + c) + np.exp(-2j * np.pi * k / 512) * (-0.5j * (me - c))
            assert np.isclose(x, X[k]), (L, a, k)
            P[k] = abs(x) ** 2
            if L == 0 and a == 0:
                P[256] = (me.real - me.imag) ** 2
    assert np.allclose(P, np.abs(X) ** 2)


def test_power_spectrum_layout_is_conflict_free_and_fits():
    for a in range(8):
        banks = {(_k(L, a) + 8 * (_k(L, a) >> 6)) % 32 for L in range(32)}
        assert len(banks) == 32, a
    pos = [k + 8 * (k >> 6) for k in range(257)]
    assert len(set(pos)) == 257 and max(pos) < 296                               # ZPAD


def _mel_tasks(lo, cnt, off):
    """Python restatement of the task builder in iss_sidekit_upload_tables."""
    cap = 1
    while sum((c + cap - 1) // cap for c in cnt) > 32:
        cap += 1
    tasks, filt = [], []
    for m in range(len(cnt)):
        parts = (cnt[m] + cap - 1) // cap
        filt.append((len(tasks), parts))
        done = 0
        for q in range(parts):
            ln = (cnt[m] - done + (parts - q) - 1) // (parts - q)
            tasks.append((lo[m] + done, ln, off[m] + done))
            done += ln
        assert done == cnt[m]
    return cap, tasks, filt


def test_mel_tasks_cover_every_filter_once():
    fb = sm.trfbank_htk24()
    fb = fb[0] if isinstance(fb, tuple) else fb
    assert fb.shape == (24, 257)
    lo, cnt, off, nnz = [], [], [], 0
    for m in range(24):
        nz = np.flatnonzero(fb[m])
        lo.append(int(nz[0])); cnt.append(int(nz[-1] - nz[0] + 1)); off.append(nnz); nnz += cnt[-1]
    cap, tasks, filt = _mel_tasks(lo, cnt, off)
    assert len(tasks) <= 32 and cap <= 24 and max(t[1] for t in tasks) <= cap
    w = np.concatenate([fb[m, lo[m]:lo[m] + cnt[m]] for m in range(24)])
    p = np.random.default_rng(1).random(257)
    for m, (t0, tn) in enumerate(filt):
        got = sum(float(np.dot(p[a:a + n], w[o:o + n])) for a, n, o in tasks[t0:t0 + tn])
        assert np.isclose(got, float(np.dot(p, fb[m])), rtol=1e-12)
