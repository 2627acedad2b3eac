"""CPU, 2 processes over gloo: the host logic of the data-parallel path (batch sharding, flat gradient buckets,
bucketed all-reduce with arm/notify/finish overlap protocol, 1/world folding) — training/train_encoder.py:115-118,185."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, dtype_name, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from omnibiote_b200.parallel import FlatGradBuckets, shard_batch
    dtype = getattr(torch, dtype_name)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(n, dtype=dtype)) for n in (100, 37, 4096, 9)]
    buckets = FlatGradBuckets([[params[0], params[1]], [params[2]], [params[3]]])
    for i, p in enumerate(params):
        assert p.grad.data_ptr() >= buckets.flat.data_ptr()
        p.grad.fill_(float(rank + 1) * (i + 1))
    # overlap protocol: bucket 1 and 0 become final during "backward", bucket 2 only at finish()
    buckets.arm()
    buckets.notify([params[2]])
    buckets.notify([params[0]])
    buckets.notify([params[1]])
    buckets.finish()
    want = sum(r + 1 for r in range(world))
    ok = all(torch.all(p.grad.float() == want * (i + 1)) for i, p in enumerate(params))
    # un-armed path (single all-reduce of everything at the end)
    for i, p in enumerate(params):
        p.grad.fill_(float(rank))
    buckets.finish()
    ok = ok and all(torch.all(p.grad.float() == sum(range(world))) for p in params)
    s, e = shard_batch(1024, world, rank)
    ok = ok and (e - s == 1024 // world) and s == rank * (1024 // world)
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name", ["float32", "bfloat16"])
def test_flat_buckets_allreduce_two_ranks(dtype_name):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), dtype_name, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_model_buckets_cover_every_parameter_once():
    from omnibiote_b200.model import OmniBioTA, OmniBioTAConfig
    from omnibiote_b200.parallel import model_buckets
    cfg = OmniBioTAConfig()
    cfg.n_layer, cfg.n_embd, cfg.n_head, cfg.vocab_size, cfg.block_size = 3, 64, 2, 128, 32
    m = OmniBioTA(cfg)
    ids = [id(p) for b in model_buckets(m) for p in b]
    assert sorted(ids) == sorted(id(p) for p in m.parameters())
    assert len(set(ids)) == len(ids)
