"""CPU, world_size 2 over gloo: the class-sharded / data-parallel step reproduces
the single-process result -- mean CE over the global batch, summed prompt gradients
(nn.DataParallel semantics of trainers/mudpt.py:230-233, :249-250)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mudpt_oracle as orc
from tests import fake_engine, golden_util as gu
from mudpt_b200 import dist as mdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, nimg, mode, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        c = gu.load(name)
        model, _ = gu.build_model(c, "cpu")
        if rank != 0:
            # replicas that start from different prompt values must be re-synchronised from rank 0
            with torch.no_grad():
                for p in model.parameters():
                    if p.requires_grad:
                        p.add_(0.05 * torch.randn_like(p))
        fake_engine.attach(model, c)
        per = nimg // world
        img = c["image"][:nimg][rank * per:(rank + 1) * per]
        lab = c["labels"][:nimg][rank * per:(rank + 1) * per]
        if mode == "fused":
            loss, logits = model.forward_backward(img, lab)
        else:
            import torch.nn.functional as F
            logits = model(img)
            # local sum / global batch; ranks' losses add up to the global mean
            loss = F.cross_entropy(logits, lab, reduction="sum") / nimg
            loss.backward()
            mdist.all_reduce_grads([p for p in model.parameters() if p.requires_grad])
            loss = mdist.all_reduce_sum(loss.detach())
        res = {"loss": float(loss), "logits": logits.detach().numpy(),
               "grads": {k: p.grad.numpy() for k, p in model.named_parameters() if p.requires_grad}}
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,nimg", [("tiny_b", 2), ("tiny_d", 4)])
@pytest.mark.parametrize("mode", ["fused", "autograd"])
def test_two_rank_step_equals_single_process(name, nimg, mode):
    c = gu.load(name)
    ref = orc.forward_backward(c["sd"], c["image"][:nimg], c["tokenized"], c["labels"][:nimg])
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, name, nimg, mode, out), nprocs=2, join=True)
    assert len(out) == 2
    logits = np.concatenate([out[0]["logits"], out[1]["logits"]], axis=0)
    np.testing.assert_allclose(logits, ref["logits"].numpy(), atol=2e-4, rtol=0)
    for r in (0, 1):
        np.testing.assert_allclose(out[r]["loss"], float(ref["loss"]), rtol=1e-5, atol=1e-5)
        for k in orc.TRAINABLE:
            g = ref["grads"][k]
            if g.numel() and float(g.norm()) > 0:
                m = orc.metrics(torch.from_numpy(out[r]["grads"][k]), g)
                assert m["cos"] > 0.99999 and m["rel_l2"] < 2e-3, (k, r, m)
    # both ranks hold identical (all-reduced) gradients
    for k in out[0]["grads"]:
        np.testing.assert_array_equal(out[0]["grads"][k], out[1]["grads"][k])


def test_shard_bounds_cover_and_balance():
    for n in (1, 7, 17, 100, 1000):
        for w in (1, 2, 3, 4, 8):
            b = [mdist.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
