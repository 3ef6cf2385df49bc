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
    full track with the very kernels a single GPU uses (replicated: the chain is
    chunk-parallel inside iss_energy_viterbi, a few ms per 10 h, so replication
    costs nothing and needs no second exchange);
    (2) per-segment Viterbi of CNN posteriors: ranks all-gather the posteriors
    of their patch range (12-16 B/patch) and decode all segments, replicated.
  * every rank ends with the complete, identical segment list.

Because every kernel is evaluated on exactly the same values in the same order
as on one GPU, the sharded result is bit-identical to the single-GPU result
(default energy_mode 'replicated').  The opt-in energy_mode 'transfer' cuts the
energy chain at the rank boundaries instead (max-plus transfer matrices); it is
exact in exact arithmetic but associates the entry scores differently, so it is
not covered by the bit-identity statement.

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

    def all_gather_var(self, t, counts=None):
        """Concatenation (rank order) of per-rank tensors whose first dimension differs.  `counts` (the
        per-rank lengths, when every rank can derive them from the plan) skips the length exchange and
        its host synchronisation."""
        if self.world == 1:
            return t
        if t.is_cuda and dist.get_backend(self.group) == 'gloo':
            # test configuration (several ranks sharing one GPU): stage through the host
            return Comm('cpu', self.group).all_gather_var(t.cpu(), counts).to(t.device)
        if counts is None:
            n = torch.tensor([t.shape[0]], dtype=torch.int64, device=self.device)
            cl = [torch.zeros_like(n) for _ in range(self.world)]
            dist.all_gather(cl, n, group=self.group)
            counts = [int(c.item()) for c in cl]
        assert counts[self.rank] == t.shape[0], (counts, self.rank, t.shape)
        m = max(counts)
        pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
        pad[:t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad, group=self.group)
        self.bytes += pad.numel() * pad.element_size() * self.world
        return torch.cat([b[:c] for b, c in zip(bufs, counts)])


    def all_gather_fixed(self, arr):
        """All-gather of a small fixed-size float64 numpy vector -> [world, n] numpy array."""
        if self.world == 1:
            return np.asarray(arr, dtype=np.float64)[None]
        dev = self.device
        if dist.get_backend(self.group) == 'gloo':
            dev = 'cpu'
        t = torch.tensor(np.asarray(arr, dtype=np.float64), device=dev)
        bufs = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(bufs, t, group=self.group)
        self.bytes += t.numel() * 8 * self.world
        return np.stack([b.cpu().numpy() for b in bufs])


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

    # -- partial energy chains (scalable variant) --
    def loge_stats(self, loge_global):
        return self.engine.loge_stats(self.ctx, loge_global)

    def energy_transfer(self, loge_own, stats, ratio):
        return self.engine.energy_transfer(self.ctx, loge_own, stats, ratio)

    def energy_forward(self, loge_own, stats, ratio, vin):
        return self.engine.energy_forward(self.ctx, loge_own, stats, ratio, vin)

    def energy_emit(self, loge_own, end_state):
        return self.engine.energy_emit(self.ctx, loge_own, end_state, out_stride=2)

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
    counts = [sum(b - a for a, b in plan.local_ranges(r, lseg, spec.inlabel)) for r in range(comm.world)]
    probs = comm.all_gather_var(probs_loc, counts)
    sel = [(a, b) for lab, a, b in lseg if lab == spec.inlabel]
    if not sel:
        return list(lseg)
    seg_off = np.concatenate(([0], np.cumsum([b - a for a, b in sel]))).astype(np.int64)
    assert probs.shape[0] == seg_off[-1], (probs.shape, seg_off[-1])
    trans = diag_trans_exp(spec.viterbi_arg, len(spec.outlabels))
    if comm.world == 1 or not getattr(plan, 'shard_decode', True):
        states = backend.viterbi(probs, seg_off, trans)
    else:
        # segments are independent chains: a rank decodes those that START in its patch range (the
        # posteriors of straddling segments are already here) and the label tracks are all-gathered
        pa, pb = plan.patch_range(rank)
        mine = [k for k, (a, b) in enumerate(sel) if pa <= a < pb]
        if mine:
            k0, k1 = mine[0], mine[-1] + 1                       # contiguous because segments are time-ordered
            local = backend.viterbi(probs[seg_off[k0]:seg_off[k1]], seg_off[k0:k1 + 1] - seg_off[k0], trans)
        else:
            local = np.zeros(0, dtype=np.uint8)
        dev = probs.device if isinstance(probs, torch.Tensor) else 'cpu'
        states = comm.all_gather_var(torch.from_numpy(np.ascontiguousarray(local, dtype=np.uint8)).to(dev)).cpu().numpy()
        assert len(states) == seg_off[-1]
    out, k = [], 0
    for lab, a, b in lseg:
        if lab != spec.inlabel:
            out.append((lab, a, b))
            continue
        for lab2, a2, b2 in _rle(states[seg_off[k]:seg_off[k + 1]]):
            out.append((spec.outlabels[lab2], a2 + a, b2 + a))
        k += 1
    return out


def _maxplus(M, v):
    """(M (x) v)[j] = max_i (M[j, i] + v[i]) in float64."""
    return np.array([max(M[j, 0] + v[0], M[j, 1] + v[1]) for j in range(2)], dtype=np.float64)


def _energy_track_transfer(backend, comm, rank, loge_own, loge_global, ratio):
    """Whole-file energy Viterbi (pyannote_viterbi.py:202-220) cut at rank boundaries, SURVEY 8(e):
    rank 0 decodes its frames from the sequence start while every other rank computes the max-plus
    transfer matrix of its frames (two basis chains); one tiny all-gather; entry scores are composed
    on the host; every rank then runs its true forward pass; a second tiny all-gather of composite
    back-pointer maps + the final argmax gives each rank its end state; tracks are all-gathered.
    Exact in exact arithmetic; floating-point association differs from the single chain only in how
    the entry scores are rounded (decisions can differ only on ~1e-16-relative near-ties)."""
    world = comm.world
    stats = backend.loge_stats(loge_global)                  # same kernel & order as on one GPU: identical threshold
    if rank == 0:
        vout, bmap = backend.energy_forward(loge_own, stats, ratio, None)
        msg = np.concatenate((np.zeros(4), vout))
    else:
        msg = np.concatenate((backend.energy_transfer(loge_own, stats, ratio).ravel(), np.zeros(2)))
    allm = comm.all_gather_fixed(msg)                         # [world, 6]
    v = allm[0, 4:6]                                          # scores leaving rank 0
    vins = [None, v]
    for r in range(1, world - 1):
        v = _maxplus(allm[r, :4].reshape(2, 2), v)
        vins.append(v)
    if rank > 0:
        vout, bmap = backend.energy_forward(loge_own, stats, ratio, vins[rank])
    tail = np.array([float(bmap[0]), float(bmap[1]), float(np.argmax(vout))])      # first max on ties, like numpy
    allt = comm.all_gather_fixed(tail)                        # [world, 3]
    end = [0] * world
    end[world - 1] = int(allt[world - 1, 2])
    for r in range(world - 2, -1, -1):
        end[r] = int(allt[r + 1, end[r + 1]])
    local = backend.energy_emit(loge_own, end[rank])
    if not isinstance(local, torch.Tensor):
        local = torch.from_numpy(np.ascontiguousarray(local, dtype=np.uint8))
    return comm.all_gather_var(local).cpu().numpy()


def segment_sharded(backend, comm, plan, pcm_local, vad_spec, gender_spec=None, energy_ratio=0.03, start_sec=0):
    """Segment the recording described by `plan`; `pcm_local` is this rank's
    sample_range().  Returns the full segment list (identical on every rank)."""
    rank = comm.rank
    fa, fb = plan.frame_range(rank)
    mspec, loge = backend.features(pcm_local)
    assert len(loge) == fb - fa, (len(loge), fa, fb)
    oa, ob = plan.owned_frames(rank)
    loge_own = loge[oa - fa:ob - fa].contiguous()
    loge_global = comm.all_gather_var(loge_own, [plan.owned_frames(r)[1] - plan.owned_frames(r)[0] for r in range(comm.world)])
    assert loge_global.shape[0] == plan.L
    mode = getattr(plan, 'energy_mode', 'auto')
    if mode == 'auto':
        mode = 'replicated'
    if comm.world == 1 or mode == 'replicated':
        track = backend.energy_track(loge_global, energy_ratio)
    else:
        track = _energy_track_transfer(backend, comm, rank, loge_own, loge_global, energy_ratio)
        assert len(track) == plan.P, (len(track), plan.P)
    lseg = [('noEnergy' if lab == 0 else 'energy', a, b) for lab, a, b in _rle(track)]
    lseg = _dnn_stage(backend, comm, plan, rank, 'vad', vad_spec, mspec, lseg)
    if gender_spec is not None:
        lseg = _dnn_stage(backend, comm, plan, rank, 'gender', gender_spec, mspec, lseg)
    return [(lab, start_sec + a * .02, start_sec + b * .02) for lab, a, b in lseg]


def segment_signal_sharded(segmenter, pcm_local, n_samples_total, group=None, energy_mode='auto'):
    """Convenience wrapper for the product: `segmenter` is this rank's Segmenter.
    energy_mode: 'replicated' (= 'auto': bit-identical to one GPU) | 'transfer' (chain cut at rank boundaries)."""
    comm = Comm(segmenter.ctx.device, group)
    plan = ShardPlan(n_samples_total, comm.world)
    plan.energy_mode = energy_mode
    backend = CudaBackend(segmenter)
    gender = segmenter.gender if segmenter.detect_gender else None
    return segment_sharded(backend, comm, plan, pcm_local, segmenter.vad, gender, segmenter.energy_ratio), comm
