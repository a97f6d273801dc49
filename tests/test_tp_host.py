"""Host-side tensor-parallel logic on CPU with the gloo backend, world_size 2: the weight slicing of
llava_align_b200.engine._tp_slice (column-parallel q/k/v/gate/up + vocab-split lm_head, row-parallel o/down) reproduces the
unsharded layer when the row-parallel partials are all-reduced -- the exchange the CUDA engine performs with NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _layer(x, w, heads):
    """One LLaMA block without norms/rope (they are replicated, not sharded), fp64."""
    H = x.shape[-1]
    q, k, v = x @ w["q"].T, x @ w["k"].T, x @ w["v"].T
    d = w["q"].shape[0] // heads
    outs = []
    for h in range(heads):
        sl = slice(h * d, (h + 1) * d)
        a = torch.softmax(q[:, sl] @ k[:, sl].T / d ** 0.5, -1)
        outs.append(a @ v[:, sl])
    attn = torch.cat(outs, -1)
    o_partial = attn @ w["o"].T
    return o_partial, lambda y: (torch.nn.functional.silu(y @ w["gate"].T) * (y @ w["up"].T)) @ w["down"].T


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from llava_align_b200.engine import _tp_slice
    torch.manual_seed(0)
    H, I, V, heads, T = 64, 96, 128, 4, 5
    full = {"q": torch.randn(H, H), "k": torch.randn(H, H), "v": torch.randn(H, H), "o": torch.randn(H, H),
            "gate": torch.randn(I, H), "up": torch.randn(I, H), "down": torch.randn(H, I), "lm": torch.randn(V, H)}
    full = {k: v.double() / 8 for k, v in full.items()}
    names = {"q": "model.layers.0.self_attn.q_proj.weight", "k": "model.layers.0.self_attn.k_proj.weight",
             "v": "model.layers.0.self_attn.v_proj.weight", "o": "model.layers.0.self_attn.o_proj.weight",
             "gate": "model.layers.0.mlp.gate_proj.weight", "up": "model.layers.0.mlp.up_proj.weight",
             "down": "model.layers.0.mlp.down_proj.weight", "lm": "lm_head.weight"}
    x = torch.randn(T, H).double()
    # unsharded
    o, mlp = _layer(x, full, heads)
    y = x + o
    y = y + mlp(y)
    logits = y @ full["lm"].T
    # sharded: same code on the rank's slices, all-reduce after the row-parallel projections, all-gather the vocab shards
    loc = {k: _tp_slice(names[k], v, world, rank) for k, v in full.items()}
    o_l, mlp_l = _layer(x, loc, heads // world)
    dist.all_reduce(o_l)
    y_l = x + o_l
    m_l = mlp_l(y_l)
    dist.all_reduce(m_l)
    y_l = y_l + m_l
    shard = y_l @ loc["lm"].T
    parts = [torch.empty_like(shard) for _ in range(world)]
    dist.all_gather(parts, shard)
    got = torch.cat(parts, -1)
    ok = torch.allclose(got, logits, rtol=1e-9, atol=1e-9) and loc["o"].shape == (H, H // world) and loc["lm"].shape == (V // world, H)
    # unique-id style plumbing: rank 0's bytes reach every rank
    idt = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(idt, 0)
    ok = ok and bytes(idt.numpy().tobytes()) == bytes(range(128))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_tp_slicing_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
