"""CPU: the oracle restatement against the committed reference goldens
(tests/golden/reference_golden.npz, produced by tests/golden/make_golden.py
from the real reference modules) and the reference's own golden CSVs."""
import hashlib
import os

import numpy as np
import pytest

from oracle import segmenter_oracle as so
from oracle import sidekit_oracle as sk
from oracle import viterbi_oracle as vo


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_filterbank_table(golden):
    fb = sk.mel_filterbank()[0]
    assert fb.dtype == np.float32 and np.array_equal(fb, golden['fbank'])
    assert np.count_nonzero(fb) == 454


@pytest.mark.parametrize('name', ['musanmix', 'silence2sec', 'lamartine', 'synth'])
def test_frontend_bit_exact(golden, media, name):
    sig = golden['synth_sig'] if name == 'synth' else so.read_wav_16k_mono(os.path.join(media, name + '.wav'))
    assert len(sig) == int(golden[name + '_nsamp'])
    mspec, loge = sk.logmel_loge(sig)
    rows = golden[name + '_rows']
    assert np.array_equal(mspec[rows], golden[name + '_mspec'], equal_nan=True)
    assert np.array_equal(loge[rows], golden[name + '_loge'], equal_nan=True)
    assert _sha(mspec) == str(golden[name + '_sha_mspec'])
    assert _sha(loge) == str(golden[name + '_sha_loge'])


def test_viterbi_cases(golden):
    for i in range(int(golden['vit_ncases'])):
        em, tr, st = golden['vit%d_em' % i], golden['vit%d_tr' % i], golden['vit%d_st' % i]
        assert np.array_equal(vo.viterbi_c(em, tr), st.astype(np.float64))
        if len(em) <= 2000:
            assert np.array_equal(vo.viterbi_numpy(em, tr), st.astype(np.float64))


def _csv_rows(path):
    rows = []
    with open(path) as f:
        next(f)
        for line in f:
            lab, a, b = line.rstrip('\n').split('\t')
            rows.append((lab, float(a), float(b)))
    return rows


@pytest.mark.parametrize('csv', ['musanmix-smn-gender.csv', 'musanmix-sm-gender.csv'])
def test_musanmix_noenergy_rows(media, csv):
    """Weights-free known answer: every noEnergy row (and every boundary) of the
    reference golden CSVs is reproduced exactly (same doubles)."""
    sig = so.read_wav_16k_mono(os.path.join(media, 'musanmix.wav'))
    mspec, loge, difflen = so.media2feats(sig)
    segs = [(lab, a * .02, b * .02) for lab, a, b in so.energy_segments(loge)]
    ref = _csv_rows(os.path.join(media, csv))
    assert [s for s in segs if s[0] == 'noEnergy'] == [r for r in ref if r[0] == 'noEnergy']
    # every energy/noEnergy boundary is a boundary of the golden segmentation
    ref_bounds = {r[1] for r in ref} | {ref[-1][2]}
    assert {s[1] for s in segs} <= ref_bounds and segs[-1][2] == ref[-1][2]


def test_silence_known_answer(media):
    sig = so.read_wav_16k_mono(os.path.join(media, 'silence2sec.wav'))
    mspec, loge, difflen = so.media2feats(sig)
    segs = [(lab, a * .02, b * .02) for lab, a, b in so.energy_segments(loge)]
    assert segs == _csv_rows(os.path.join(media, 'silence2sec-smn-gender.csv')) == [('noEnergy', 0.0, 1.98)]


def test_get_patches_shape_and_replication():
    rng = np.random.default_rng(0)
    for L in (68, 69, 100, 101, 333):
        m = rng.standard_normal((L, 21)).astype(np.float32)
        p, fin = so.get_patches(m)
        assert len(p) == (L + 1) // 2 and fin.all()          # exactly ceil(L/2) patches
        assert np.array_equal(p[0], p[17]) and np.array_equal(p[-1], p[-(16 + L % 2) - 1])
    m[40, 3] = -np.inf
    p, fin = so.get_patches(m)
    assert not fin.all() and fin.any()
