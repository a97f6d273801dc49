"""Tensor-parallel engine on 2 GPUs (NCCL over NVLink) == the single-GPU engine: greedy VDD tokens identical, stream logits
within fp32 re-association noise.  Needs >= 2 CUDA devices (run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret, B=2, use_none=False):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from helpers import make_inputs, tiny_model
    from llava_align_b200 import engine as E
    from llava_align_b200.model import LlavaLlamaForCausalLM, evolve_vcd_sampling
    cfg, sd, oracle = tiny_model(0)
    evolve_vcd_sampling()
    m = LlavaLlamaForCausalLM(cfg, f"cuda:{rank}", tp_size=world, tp_rank=rank)
    m.load_state_dict(sd)
    m.engine.tp_setup(dist)
    ids, img = make_inputs(cfg, B, 12, 9)
    n = 6
    m.engine.encode_images(img)
    lg = m.engine.prefill(ids, True, True, use_none, n).cpu().clone()
    out = m.generate(ids.cuda(rank), images=img.cuda(rank), use_dd_unk=True, use_dd=use_none, cd_alpha=1.0, cd_beta=0.1, do_sample=True,
                     top_k=1, max_new_tokens=n, eos_token_id=None)
    ret[rank] = (lg, out.cpu())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_tp2_equals_tp1():
    from helpers import make_inputs, tiny_model
    from llava_align_b200.model import LlavaLlamaForCausalLM, evolve_vcd_sampling
    cfg, sd, oracle = tiny_model(0)
    evolve_vcd_sampling()
    m = LlavaLlamaForCausalLM(cfg, "cuda:0")
    m.load_state_dict(sd)
    ids, img = make_inputs(cfg, 2, 12, 9)
    n = 6
    m.engine.encode_images(img)
    lg1 = m.engine.prefill(ids, True, True, False, n).cpu().clone()
    out1 = m.generate(ids.cuda(), images=img.cuda(), use_dd_unk=True, cd_alpha=1.0, cd_beta=0.1, do_sample=True, top_k=1,
                      max_new_tokens=n, eos_token_id=None).cpu()
    del m
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for r in range(2):
        lg, out = ret[r]
        assert (lg - lg1).abs().max().item() < 5e-2  # K split in two halves: fp32 partial sums re-associated
        assert torch.equal(out, out1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_tp2_many_rows_equals_tp1():
    """15 fused rows (batch 5, three streams): the tensor-core decode path with NCCL all-reduces and the streaming attention
    kernel over the local heads, against the single-GPU engine."""
    from helpers import make_inputs, tiny_model
    from llava_align_b200.model import LlavaLlamaForCausalLM, evolve_vcd_sampling
    cfg, sd, oracle = tiny_model(0)
    evolve_vcd_sampling()
    m = LlavaLlamaForCausalLM(cfg, "cuda:0")
    m.load_state_dict(sd)
    ids, img = make_inputs(cfg, 5, 12, 9)
    n = 6
    m.engine.encode_images(img)
    lg1 = m.engine.prefill(ids, True, True, True, n).cpu().clone()
    out1 = m.generate(ids.cuda(), images=img.cuda(), use_dd_unk=True, use_dd=True, cd_alpha=1.0, cd_beta=0.1, do_sample=True, top_k=1,
                      max_new_tokens=n, eos_token_id=None).cpu()
    del m
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret, 5, True), nprocs=2, join=True)
    for r in range(2):
        lg, out = ret[r]
        assert (lg - lg1).abs().max().item() < 5e-2
        if not torch.equal(out, out1):  # a greedy token may only differ where the single-GPU top-2 margin is inside the noise
            first = int((out != out1).nonzero()[0, 1])
            assert first >= 12
