"""Tensor-parallel engine on N GPUs (peer-memory all-reduce inside the decode kernels, NCCL for the GEMM-sized paths) == the
single-GPU engine: greedy VDD tokens identical unless the single-GPU run itself sits at a near-tie, stream logits within
fp32 re-association noise, sampled tokens identical on every rank.  Needs >= 2 CUDA devices (`gpurun --gpus N`); the
real-width cases run at every N in {2, 4, 8} the box offers."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

LOGIT_TOL_TP = 5e-2   # K split across ranks: fp32 partial sums re-associated, bf16 activations re-rounded
MARGIN_TOL = 0.25     # combined scores amplify a logit difference by (1 + alpha) + alpha = 3 at alpha = 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(model, rank, world):
    """model: 'tiny' (exact CPU-reproducible weights) or a by_name() spec such as 'llava-v1.5-7b@4' (on-device seeded init,
    identical on every rank: the same generator seed draws the same full tensor, each rank keeps its slice)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from llava_align_b200.model import LlavaLlamaForCausalLM, evolve_vcd_sampling
    evolve_vcd_sampling()
    if model == "tiny":
        from helpers import tiny_model
        cfg, sd, _ = tiny_model(0)
        m = LlavaLlamaForCausalLM(cfg, f"cuda:{rank}", tp_size=world, tp_rank=rank)
        m.load_state_dict(sd)
        return cfg, m
    from llava_align_b200.builder import load_pretrained_model
    _, m, _, _ = load_pretrained_model(f"random:{model}:0", None, model, device=f"cuda:{rank}", tp_size=world, tp_rank=rank)
    return m.vdd_config, m


def _run(m, cfg, rank, B, use_none, n, sampler, seed):
    """Returns (new tokens [B, n], per-step stream logits [n, S, B, V] teacher-forced on those tokens)."""
    from helpers import make_inputs
    ids, img = make_inputs(cfg, B, 12, 9)
    torch.manual_seed(seed)  # TP ranks pass DIFFERENT seeds: generate() must make the variates agree by itself
    out = m.generate(ids.cuda(rank), images=img.cuda(rank), use_dd_unk=True, use_dd=use_none, cd_alpha=1.0, cd_beta=0.1,
                     do_sample=True, max_new_tokens=n, eos_token_id=None, **sampler)
    toks = out[:, ids.shape[1]:].cpu()
    m.engine.encode_images(img)
    lg = m.engine.prefill(ids, True, True, use_none, n)
    steps = []
    for st in range(n):
        steps.append(lg.cpu().clone())
        if st + 1 < n:
            lg = m.engine.decode_step(toks[:, st])
    return toks, torch.stack(steps)


def _worker(rank, world, port, ret, model, B, use_none, n, sampler):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg, m = _build(model, rank, world)
    m.engine.tp_setup(dist)
    ret[rank] = _run(m, cfg, rank, B, use_none, n, sampler, seed=1000 * rank)
    dist.barrier()
    dist.destroy_process_group()


def _compare(world, model, B, use_none, n=6, sampler=None, logit_tol=LOGIT_TOL_TP):
    from oracle.vdd_oracle import contrastive_combine  # test infrastructure: the combine the greedy decision is taken on
    sampler = dict(top_k=1) if sampler is None else sampler
    cfg, m = _build(model, 0, 1)
    toks1, lg1 = _run(m, cfg, 0, B, use_none, n, sampler, seed=0)
    del m
    torch.cuda.empty_cache()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret, model, B, use_none, n, sampler), nprocs=world, join=True)
    for r in range(world):
        toks, lg = ret[r]
        assert torch.equal(toks, ret[0][0]), f"rank {r} emitted different tokens than rank 0"
        # step 0 (prefill) is free of feedback; later steps are comparable for the rows whose tokens still agree
        same = torch.ones(B, dtype=torch.bool)
        for st in range(n):
            err = (lg[st][:, same] - lg1[st][:, same]).abs().max().item() if bool(same.any()) else 0.0
            assert err < logit_tol, (r, st, err)
            if sampler.get("top_k") == 1:
                l = lg1[st][0]
                cd = lg1[st][1:].mean(dim=0)  # (u + d) / 2 with both streams, u alone otherwise
                s = contrastive_combine(l, cd, 1.0, 0.1)
                for b in (toks[:, st] != toks1[:, st]).nonzero()[:, 0].tolist():
                    if same[b]:  # first divergence of this row: only legitimate at a near-tie of the single-GPU run
                        gap = (s[b].max() - s[b, toks[b, st]]).item()
                        assert gap < MARGIN_TOL, (r, st, b, gap)
            same &= toks[:, st] == toks1[:, st]
    return ret, (toks1, lg1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_tp2_equals_tp1():
    """6 fused rows (batch 2 x {main, unk, none}... here main + unk): the GEMV chain with the in-kernel all-reduce."""
    _compare(2, "tiny", 2, False)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_tp2_many_rows_equals_tp1():
    """15 fused rows (batch 5, three streams): the tensor-core decode path + streaming attention over the local heads."""
    _compare(2, "tiny", 5, True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_tp2_sampling_ranks_agree():
    """Real sampling (top_k != 1) with every rank's torch generator seeded DIFFERENTLY: the variates are broadcast from rank 0,
    so all ranks draw the same tokens (their sharded KV caches would silently diverge otherwise) -- and rank 0's variates are
    the ones a single-GPU run with the same seed draws, so the tokens match that run wherever its race is not a near-tie."""
    ret, (toks1, lg1) = _compare(2, "tiny", 2, False, n=8, sampler=dict(top_k=None, top_p=0.9, temperature=0.8))
    assert ret[0][0].shape == toks1.shape


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("B,use_none", [(1, True), (4, True)])
def test_tp_real_width_equals_tp1(world, B, use_none):
    """LLaVA-1.5-7B widths (H=4096, I=11008, V=32000, 32 heads) with 4 decoder layers: R = 3 rows (GEMV chain, in-kernel
    all-reduce) and R = 12 rows (tensor-core path), TP = 2/4/8 against TP = 1."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _compare(world, "llava-v1.5-7b@4", B, use_none, n=8, logit_tol=8e-2)
