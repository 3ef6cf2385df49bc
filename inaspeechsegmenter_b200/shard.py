"""Time-sharding of ONE long recording across GPUs.

The reference has no intra-file parallelism (its only distribution is a Pyro4
farm handing out whole files, scripts/ina_speech_segmenter_pyro_*.py).  Here a
recording is cut into contiguous patch ranges, one per rank (one process per
GPU, torch.distributed over NCCL/NVLink as plumbing):

  * each rank runs K1 on its own samples plus a halo of 34 frames (the 68-frame
    patch centred on a cut) -- the expensive per-frame work (features, both
    CNNs) is embarrassingly parallel and never leaves the rank;
  * the two global couplings of the algorithm are resolved with tiny
    exchanges: (1) the energy threshold is the mean of loge over the WHOLE file
    (segmenter.py:70) and the energy Viterbi is a whole-file chain
    (pyannote_viterbi.py:202-220): ranks all-gather their owned loge
    (4 B/frame) and every rank evaluates the reduction + the K = 2 chain on the
    full track (replicated: it is a serial ~12 ns/frame chain, so this costs the
    same as computing it once, and needs no second exchange);
    (2) per-segment Viterbi of CNN posteriors: ranks all-gather the posteriors
    of their patch range (12-16 B/patch) and decode all segments, replicated.
  * every rank ends with the complete, identical segment list.

Because every kernel is evaluated on exactly the same values in the same order
as on one GPU, the sharded result is bit-identical to the single-GPU result.

The numeric steps sit behind a small backend interface so the orchestration
and the collectives can be tested with world_size 2 on the gloo backend
(tests/test_shard_gloo.py injects a CPU backend).
"""
import numpy as np
import torch
import torch.distributed as dist

PATCH_W, PATCH_HOP, LFILL = 68, 2, 17
WIN, HOP = 400, 160


def num_frames(n):
    return 0 if n < WIN else (n - WIN) // HOP + 1


class ShardPlan:
    """Index arithmetic of the cut (pure Python, testable without devices)."""

    def __init__(self, n_samples, world):
        self.n, self.world = int(n_samples), int(world)
        self.L = num_frames(self.n)
        if self.L < PATCH_W:
            raise ValueError('recording too short to shard (%d frames)' % self.L)
        self.P = (self.L + 1) // 2                       # patches == energy labels (segmenter.py:262)
        self.U = (self.L - PATCH_W) // PATCH_HOP + 1     # un-replicated windows (segmenter.py:78)
        if self.P // self.world < 4 * LFILL:
            raise ValueError('shards of %d patches are too small (need >= %d)' % (self.P // self.world, 4 * LFILL))
        self.bounds = [r * self.P // self.world for r in range(self.world + 1)]

    def patch_range(self, r):
        return self.bounds[r], self.bounds[r + 1]

    def frame_range(self, r):
        """[fa, fb): log-mel rows rank r must compute (own rows + 34-frame halo); fa is even."""
        pa, pb = self.patch_range(r)
        fa = 0 if r == 0 else 2 * (pa - LFILL)
        fb = self.L if r == self.world - 1 else 2 * (pb - 1 - LFILL) + PATCH_W
        assert 0 <= fa <= 2 * pa and min(2 * pb, self.L) <= fb <= self.L
        return fa, fb

    def sample_range(self, r):
        fa, fb = self.frame_range(r)
        return HOP * fa, HOP * (fb - 1) + WIN

    def owned_frames(self, r):
        pa, pb = self.patch_range(r)
        return 2 * pa, min(2 * pb, self.L)

    def edges(self, r):
        return r == 0, r == self.world - 1

    def local_ranges(self, r, lseg, inlabel):
        """Global padded-patch ranges of `inlabel` segments clipped to rank r,
        as LOCAL indices for iss_cnn_forward (edge flags from edges(r))."""
        pa, pb = self.patch_range(r)
        shift = 0 if r == 0 else pa               # local index = p - 17 - fa/2 = p - pa
        out = []
        for lab, a, b in lseg:
            if lab != inlabel:
                continue
            a2, b2 = max(a, pa), min(b, pb)
            if b2 > a2:
                out.append((a2 - shift, b2 - shift))
        return out


class Comm:
    """The two collectives the path needs, over torch.distributed (NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self, device, group=None):
        self.device, self.group = device, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bytes = 0

    def all_gather_var(self, t):
        """Concatenation (rank order) of per-rank tensors whose first dimension differs."""
        if self.world == 1:
            return t
        if t.is_cuda and dist.get_backend(self.group) == 'gloo':
            # test configuration (several ranks sharing one GPU): stage through the host
            return Comm('cpu', self.group).all_gather_var(t.cpu()).to(t.device)
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=self.device)
        counts = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(counts, n, group=self.group)
        counts = [int(c.item()) for c in counts]
        m = max(counts)
        pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
        pad[:t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad, group=self.group)
        self.bytes += pad.numel() * pad.element_size() * self.world
        return torch.cat([b[:c] for b, c in zip(bufs, counts)])


class CudaBackend:
    """The product backend: libiss_b200 kernels on this rank's GPU."""

    def __init__(self, segmenter):
        from . import engine
        from .segmenter import _get_frontend
        self.seg, self.engine = segmenter, engine
        self.ctx = segmenter.ctx
        self.fe = _get_frontend('main', segmenter.device)
        self.device = self.ctx.device

    def features(self, pcm_local):
        if not isinstance(pcm_local, torch.Tensor):
            pcm_local = torch.from_numpy(np.ascontiguousarray(pcm_local))
        if not pcm_local.is_cuda:
            pcm_local = pcm_local.to(self.device, non_blocking=True)
        mspec, loge, _ = self.fe(pcm_local, self.seg.fft_precision)
        return mspec, loge

    def energy_track(self, loge_global, ratio):
        stats = self.engine.loge_stats(self.ctx, loge_global)
        return self.engine.energy_viterbi(self.ctx, loge_global, stats, ratio, out_stride=2).cpu().numpy()

    def cnn_probs(self, which, mspec_local, ranges, edge_left, edge_right):
        net = self.seg.vad.nn if which == 'vad' else self.seg.gender.nn
        return net.forward(mspec_local, ranges, edge_left, edge_right)

    def viterbi(self, probs_global, seg_off, trans):
        return self.engine.viterbi_segments(self.ctx, probs_global, seg_off, trans).cpu().numpy()


def _rle(track):
    cut = np.flatnonzero(track[1:] != track[:-1]) + 1
    starts = np.concatenate(([0], cut))
    stops = np.concatenate((cut, [len(track)]))
    return [(int(track[a]), int(a), int(b)) for a, b in zip(starts, stops)]


def _dnn_stage(backend, comm, plan, rank, which, spec, mspec_local, lseg):
    """One DnnSegmenter.__call__ (segmenter.py:135-179), CNN sharded, Viterbi replicated."""
    from .engine import diag_trans_exp
    el, er = plan.edges(rank)
    ranges = plan.local_ranges(rank, lseg, spec.inlabel)
    probs_loc = backend.cnn_probs(which, mspec_local, ranges, el, er)
    probs = comm.all_gather_var(probs_loc)
    sel = [(a, b) for lab, a, b in lseg if lab == spec.inlabel]
    if not sel:
        return list(lseg)
    seg_off = np.concatenate(([0], np.cumsum([b - a for a, b in sel]))).astype(np.int64)
    assert probs.shape[0] == seg_off[-1], (probs.shape, seg_off[-1])
    states = backend.viterbi(probs, seg_off, diag_trans_exp(spec.viterbi_arg, len(spec.outlabels)))
    out, k = [], 0
    for lab, a, b in lseg:
        if lab != spec.inlabel:
            out.append((lab, a, b))
            continue
        for lab2, a2, b2 in _rle(states[seg_off[k]:seg_off[k + 1]]):
            out.append((spec.outlabels[lab2], a2 + a, b2 + a))
        k += 1
    return out


def segment_sharded(backend, comm, plan, pcm_local, vad_spec, gender_spec=None, energy_ratio=0.03, start_sec=0):
    """Segment the recording described by `plan`; `pcm_local` is this rank's
    sample_range().  Returns the full segment list (identical on every rank)."""
    rank = comm.rank
    fa, fb = plan.frame_range(rank)
    mspec, loge = backend.features(pcm_local)
    assert len(loge) == fb - fa, (len(loge), fa, fb)
    oa, ob = plan.owned_frames(rank)
    loge_global = comm.all_gather_var(loge[oa - fa:ob - fa].contiguous())
    assert loge_global.shape[0] == plan.L
    track = backend.energy_track(loge_global, energy_ratio)
    lseg = [('noEnergy' if lab == 0 else 'energy', a, b) for lab, a, b in _rle(track)]
    lseg = _dnn_stage(backend, comm, plan, rank, 'vad', vad_spec, mspec, lseg)
    if gender_spec is not None:
        lseg = _dnn_stage(backend, comm, plan, rank, 'gender', gender_spec, mspec, lseg)
    return [(lab, start_sec + a * .02, start_sec + b * .02) for lab, a, b in lseg]


def segment_signal_sharded(segmenter, pcm_local, n_samples_total, group=None):
    """Convenience wrapper for the product: `segmenter` is this rank's Segmenter."""
    comm = Comm(segmenter.ctx.device, group)
    plan = ShardPlan(n_samples_total, comm.world)
    backend = CudaBackend(segmenter)
    gender = segmenter.gender if segmenter.detect_gender else None
    return segment_sharded(backend, comm, plan, pcm_local, segmenter.vad, gender, segmenter.energy_ratio), comm
