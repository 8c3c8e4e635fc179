"""CPU, world_size 2, gloo: the N>1 path of the KD step -- utterance sharding and the single gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from distil_whisper_b200 import ddp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from distil_whisper_b200.modeling import DistilWhisperB200ForConditionalGeneration
    from distil_whisper_b200.optim import FlatBuffers, decay_split
    from oracle import whisper_oracle as wo
    r, _, w = ddp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                       # different init per rank on purpose
    m = DistilWhisperB200ForConditionalGeneration(wo.PRESETS["tiny-student"].to_dict())
    for p in m.model.encoder.parameters():
        p.requires_grad = False
    ddp.broadcast_parameters(m, 0)
    fb = FlatBuffers(list(decay_split(m)))
    ref = [p.detach().clone() for p in m.parameters()]
    gathered = [None] * world
    dist.all_gather_object(gathered, float(sum(p.double().sum() for p in ref)))
    assert abs(gathered[0] - gathered[1]) < 1e-9            # same weights everywhere after the broadcast
    # rank-dependent gradients -> one all-reduce -> sum; averaging is the optimiser's grad_scale
    for i, (p, _) in enumerate(fb.layout):
        p.grad.fill_(float(rank + 1) * (i + 1))
    n = fb.all_reduce()
    assert n == world
    for i, (p, _) in enumerate(fb.layout):
        assert torch.all(p.grad == 3.0 * (i + 1))
    batch = wo.synthetic_batch(wo.PRESETS["tiny-student"], batch=5, n_tok=6, seed=1)
    local = ddp.shard_batch(batch, rank, world)
    sizes = [None] * world
    dist.all_gather_object(sizes, local["labels"].shape[0])
    assert sum(sizes) == 5 and max(sizes) - min(sizes) <= 1
    assert abs(ddp.max_over_ranks(float(rank + 1)) - 2.0) < 1e-12
    ret[rank] = True
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_allreduce_and_sharding():
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_shard_range_partitions():
    for n in (1, 5, 32, 33, 256):
        for w in (1, 2, 3, 8):
            spans = [ddp.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
