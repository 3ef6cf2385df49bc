"""N > 1 path on CPU: world_size-2 gloo run of the time-sharding orchestration
(inaspeechsegmenter_b200/shard.py) with a CPU backend built from the oracle.
Checks the cut arithmetic (halo, edge replication, local patch indices) and the
two collectives: the sharded segmentation must equal the unsharded one."""
import os
import socket
import sys
import warnings

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import synth_audio
from inaspeechsegmenter_b200 import models
from inaspeechsegmenter_b200.shard import Comm, ShardPlan, segment_sharded


class Spec:
    def __init__(self, d):
        self.inlabel, self.outlabels, self.viterbi_arg, self.nmel = d['inlabel'], d['outlabels'], d['viterbi_arg'], d['nmel']


class OracleBackend:
    """CPU stand-in for CudaBackend with the addressing semantics of iss_cnn_forward."""

    def __init__(self, mods):
        from oracle import cnn_oracle
        self.nets = {'vad': (cnn_oracle.KerasLikeModel(*mods['vad'], threads=2), 21),
                     'gender': (cnn_oracle.KerasLikeModel(*mods['gender'], threads=2), 24)}
        self.device = 'cpu'

    def features(self, pcm_local):
        from oracle import sidekit_oracle as sk
        sig = np.asarray(pcm_local).astype(np.float32) / np.float32(32768)
        mspec, loge = sk.logmel_loge(sig)
        return torch.from_numpy(mspec), torch.from_numpy(loge)

    def energy_track(self, loge_global, ratio):
        from oracle import segmenter_oracle as so
        return so.energy_activity(loge_global.numpy(), ratio)[::2].astype(np.uint8)

    # ---- partial energy chains (numpy restatement of the same DP, for the 'transfer' mode) ----
    def loge_stats(self, loge_global):
        return loge_global.numpy()

    def _emissions(self, loge_own, loge_global_np, ratio):
        from oracle import viterbi_oracle as vo
        g = loge_global_np
        thr = np.mean(g[np.isfinite(g)]) + np.log(ratio)                # segmenter.py:70
        return vo.pred2logemission(loge_own.numpy() > thr), vo.log_trans_exp(150, cost0=-5)

    def _run(self, em, A, V, first_is_init):
        bps = np.zeros((len(em), 2), dtype=np.int64)
        for t in range(len(em)):
            if t == 0 and first_is_init:
                V = em[0] + np.log(np.ones(2) / 2)
                bps[0] = (0, 1)
            else:
                cand = V[:, None] + A
                bps[t] = np.argmax(cand, axis=0)
                V = em[t] + cand[bps[t], np.arange(2)]
        return V, bps

    def energy_transfer(self, loge_own, stats, ratio):
        em, A = self._emissions(loge_own, stats, ratio)
        M = np.zeros((2, 2))
        for i in range(2):
            v0 = np.full(2, -np.inf)
            v0[i] = 0.0
            M[:, i] = self._run(em, A, v0, False)[0]
        return M

    def energy_forward(self, loge_own, stats, ratio, vin):
        em, A = self._emissions(loge_own, stats, ratio)
        V, bps = self._run(em, A, None if vin is None else np.asarray(vin, dtype=np.float64), vin is None)
        self._bps, self._vout = bps, V
        bmap = np.zeros(2, dtype=np.uint8)
        for x in range(2):
            s_ = x
            for t in range(len(em) - 1, 0 if vin is None else -1, -1):
                s_ = bps[t, s_]
            bmap[x] = s_
        return V, bmap

    def energy_emit(self, loge_own, end_state):
        x = int(np.argmax(self._vout)) if end_state < 0 else int(end_state)
        T = len(self._bps)
        st = np.zeros(T, dtype=np.uint8)
        for t in range(T - 1, -1, -1):
            st[t] = x
            x = self._bps[t, x]
        return st[::2]

    def cnn_probs(self, which, mspec_local, ranges, edge_left, edge_right):
        net, nmel = self.nets[which]
        m = mspec_local.numpy()[:, :nmel]
        L = len(m)
        U = (L - 68) // 2 + 1
        idx = np.concatenate([np.arange(a, b) for a, b in ranges]) if ranges else np.zeros(0, np.int64)
        K = 3 if which == 'vad' else 2
        if len(idx) == 0:
            return torch.zeros((0, K), dtype=torch.float32)
        j = idx - (17 if edge_left else 0)
        j = np.maximum(j, 0)
        if edge_right:
            j = np.minimum(j, U - 1)
        assert j.max() <= U - 1
        win = np.lib.stride_tricks.sliding_window_view(m, (68, nmel))[::2, 0][j].reshape(len(j), -1)
        with np.errstate(invalid='ignore', divide='ignore'):
            x = (win - win.mean(1, keepdims=True)) / win.std(1, keepdims=True)
        finite = np.isfinite(x).all(1)
        p = net.predict(x.reshape(-1, 68, nmel, 1).astype(np.float32))
        p[~finite] = 0.5
        return torch.from_numpy(p)

    def viterbi(self, probs, seg_off, trans):
        from oracle import viterbi_oracle as vo
        out = np.zeros(len(probs), np.uint8)
        p = probs.numpy()
        for s in range(len(seg_off) - 1):
            a, b = seg_off[s], seg_off[s + 1]
            with np.errstate(divide='ignore'):
                out[a:b] = vo.viterbi_c(np.log(p[a:b]), trans)
        return out


def _mods():
    return {'vad': models.synthetic_keras_cnn(21, 3, seed=11, width=0.25),
            'gender': models.synthetic_keras_cnn(24, 2, seed=13, width=0.25)}


def _specs():
    from oracle import segmenter_oracle as so
    return Spec(so.VAD_SMN), Spec(so.GENDER)


def _worker(rank, world, port, s16, q, energy_mode='auto'):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        plan = ShardPlan(len(s16), world)
        plan.energy_mode = energy_mode
        sa, sb = plan.sample_range(rank)
        vad, gender = _specs()
        comm = Comm('cpu')
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            segs = segment_sharded(OracleBackend(_mods()), comm, plan, s16[sa:sb], vad, gender)
        q.put((rank, segs, comm.bytes))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_plan_arithmetic():
    for n, world in ((16000 * 60 + 123, 2), (16000 * 61, 3), (16000 * 200 + 7, 8)):
        plan = ShardPlan(n, world)
        assert plan.bounds[0] == 0 and plan.bounds[-1] == plan.P == (plan.L + 1) // 2
        cover = []
        for r in range(world):
            pa, pb = plan.patch_range(r)
            fa, fb = plan.frame_range(r)
            sa, sb = plan.sample_range(r)
            oa, ob = plan.owned_frames(r)
            assert fa % 2 == 0 and 0 <= sa < sb <= n and fa <= oa <= ob <= fb <= plan.L
            assert (sb - sa - 400) // 160 + 1 == fb - fa           # K1 yields exactly the rows needed
            # every patch of the rank reads rows inside [fa, fb)
            for p in (pa, pb - 1):
                j = min(max(p - 17, 0), plan.U - 1)
                assert fa <= 2 * j and 2 * j + 68 <= fb
            cover.append((oa, ob))
        assert cover[0][0] == 0 and cover[-1][1] == plan.L and all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        ShardPlan(16000 * 5, 8)


@pytest.mark.parametrize('world,energy_mode', [(2, 'replicated'), (2, 'transfer'), (3, 'auto')])
def test_sharded_equals_unsharded_gloo(world, energy_mode):
    from oracle import cnn_oracle, segmenter_oracle as so
    s16 = synth_audio(50, seed=21)
    mods = _mods()
    mspec, loge, difflen = so.media2feats(s16.astype(np.float32) / np.float32(32768))
    v = so.DnnSegmenterOracle(cnn_oracle.KerasLikeModel(*mods['vad'], threads=2), **so.VAD_SMN)
    g = so.DnnSegmenterOracle(cnn_oracle.KerasLikeModel(*mods['gender'], threads=2), **so.GENDER)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = so.segment_feats(mspec, loge, difflen, 0, v, g)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, s16, q, energy_mode)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, segs, nbytes in results:
        assert segs == ref, (rank, segs[:4], ref[:4])
        assert nbytes > 0
