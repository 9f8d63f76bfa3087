"""CPU, world_size 2, gloo: the host-side logic of the agent-sharded path — slot layout, the single all-gather, and
that the gathered exchange buffer reproduces the unsharded agent-major key / query / feature arrays bit for bit (so
the attention over it equals the single-process attention)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multiagentperception_b200 import sharding
from oracle import when2com_oracle as orc

AGENTS, BATCH, KD, QD, H, W, C = 4, 3, 64, 8, 2, 2, 16


def _full_arrays():
    g = torch.Generator().manual_seed(5)
    keys = torch.randn(AGENTS * BATCH, KD, generator=g)
    queries = torch.randn(AGENTS * BATCH, QD, generator=g)
    val = torch.randn(AGENTS * BATCH, H, W, C, generator=g).to(torch.bfloat16)
    return keys, queries, val


def _worker(rank, world, port, planes, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        keys, queries, val = _full_arrays()
        if planes == 2:
            val = torch.cat((val, val * 0.5), dim=-1)  # hi | lo planes
        lay = sharding.AgentShardLayout(AGENTS, world, rank, BATCH, KD, QD, H, W, C, planes)
        ex = lay.allocate("cpu")
        k, q, v = lay.views(ex)
        rows = slice(lay.first_agent * BATCH, (lay.first_agent + lay.apr) * BATCH)
        k.copy_(keys[rows])
        q.copy_(queries[rows])
        v.copy_(val[rows])
        work = sharding.all_gather_values(ex, lay, async_op=True)   # the two-step exchange of the sharded forward
        sharding.all_gather_keys_queries(ex, lay)
        work.wait()
        dk, dq, dv = lay.dense(ex)
        ok = torch.equal(dk, keys) and torch.equal(dq, queries) and torch.equal(dv, val)
        # strides the kernel will use address the same data: agent i, scene b
        # (base pointers = rank 0's arrays, as models/agents.py passes them; strides in elements of each dtype)
        flat32 = lay.kq_region(ex).view(torch.float32)
        flat16 = lay.val_region(ex).view(torch.bfloat16)
        for i in range(AGENTS):
            r, l = divmod(i, lay.apr)
            for b in range(BATCH):
                base = r * lay.keys_rank_stride + (l * BATCH + b) * KD
                ok &= torch.equal(flat32[base:base + KD], keys[i * BATCH + b])
                base = r * lay.queries_rank_stride + lay.queries_off // 4 + (l * BATCH + b) * QD
                ok &= torch.equal(flat32[base:base + QD], queries[i * BATCH + b])
                per_img = H * W * planes * C
                base = r * lay.val_rank_stride + (l * BATCH + b) * per_img
                ok &= torch.equal(flat16[base:base + per_img], val[i * BATCH + b].reshape(-1))
        # attention over the gathered arrays == attention over the unsharded ones
        sd = {"a.linear.weight": torch.randn(KD, QD, generator=torch.Generator().manual_seed(1)),
              "a.linear.bias": torch.zeros(KD)}
        def attn(kk, qq, vv):
            km = kk.view(AGENTS, BATCH, KD).transpose(0, 1)
            qm = qq.view(AGENTS, BATCH, QD).transpose(0, 1)
            vm = vv[..., :C].float().permute(0, 3, 1, 2).reshape(AGENTS, BATCH, C, H, W).transpose(0, 1)
            return orc.mimo_attention(qm, km, vm, sd, "a")
        f1, p1 = attn(dk, dq, dv)
        f2, p2 = attn(keys, queries, val)
        ok &= torch.equal(f1, f2) and torch.equal(p1, p2)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(planes):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), planes, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_all_gather_slots_world2_bf16():
    _run(1)


def test_all_gather_slots_world2_bf16x2():
    _run(2)
