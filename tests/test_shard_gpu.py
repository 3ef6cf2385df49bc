"""GPU: the time-sharded path with the real CUDA backend.  Two ranks share
cuda:0 over gloo (so it runs on a 1-GPU box); with >= 2 GPUs the same test also
runs one rank per GPU over NCCL.  The sharded result must be bit-identical to
the single-process Segmenter result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import synth_audio

pytestmark = pytest.mark.gpu


def _mods():
    from inaspeechsegmenter_b200 import models
    return {'vad': models.synthetic_keras_cnn(21, 3, seed=11, width=0.5),
            'gender': models.synthetic_keras_cnn(24, 2, seed=13, width=0.5)}


def _worker(rank, world, port, s16, backend, q, energy_mode='auto'):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dev = rank if backend == 'nccl' else 0
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        from inaspeechsegmenter_b200 import Segmenter
        from inaspeechsegmenter_b200.shard import ShardPlan, segment_signal_sharded
        seg = Segmenter(vad_engine='smn', detect_gender=True, ffmpeg=None, models=_mods(), device=dev)
        plan = ShardPlan(len(s16), world)
        sa, sb = plan.sample_range(rank)
        segs, comm = segment_signal_sharded(seg, s16[sa:sb], len(s16), energy_mode=energy_mode)
        ref = seg.segment_signal(s16) if rank == 0 else None
        q.put((rank, segs, ref, comm.bytes))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, backend, energy_mode='auto'):
    s16 = synth_audio(120, seed=33)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, s16, backend, q, energy_mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=900) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = res[0][2]
    assert len(ref) > 3
    for rank, segs, _, nbytes in res:
        assert segs == ref, (rank, segs[:4], ref[:4])


@pytest.mark.parametrize('world,energy_mode', [(2, 'replicated'), (2, 'transfer'), (3, 'transfer')])
def test_sharded_ranks_one_gpu_gloo(world, energy_mode):
    """`transfer` = the whole-file energy chain cut at rank boundaries (max-plus transfer matrices)."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    _run(world, 'gloo', energy_mode)


def test_sharded_nccl_one_rank_per_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    _run(min(torch.cuda.device_count(), 4), 'nccl')
