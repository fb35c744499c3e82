"""Multi-GPU parity check (run under torchrun on the GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/gpu_dist_check.py

Every rank runs the class-sharded / data-parallel fused step (NCCL all-gather of text features,
reduce-scatter of their gradient, all-reduce of the prompt gradients) on its slice of a global
batch; rank 0 also runs the same global batch unsharded on its own GPU and compares loss, logits
and the 10 gradients.  Writes gpurun_out/dist_check.json.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mudpt_b200 import synthetic as syn  # noqa: E402
from tests.test_gpu_parity import _build_big  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    B, C = 8, 250  # per-rank batch, total classes (250 / 4 ranks is uneven: exercises the padded path at world 4)
    model = _build_big(B, C).to(dev)
    images = syn.synthetic_images(B * world, 224, seed=7).to(dev)
    labels = syn.synthetic_labels(B * world, C, seed=7).to(dev)
    model.zero_grad(set_to_none=True)
    loss, logits = model.forward_backward(images[rank * B:(rank + 1) * B], labels[rank * B:(rank + 1) * B])
    torch.cuda.synchronize()
    grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad}
    gathered = [torch.empty_like(logits) for _ in range(world)]
    dist.all_gather(gathered, logits)
    out = {"world": world}
    # diagnostics: sharded text features (gathered) and image features of this rank
    with torch.no_grad():
        eng = model._clip_ref[0].engine()
        P_v, P_t = model.prompt_stacks()
        ft_loc = eng.text_forward(P_t, True)
        ft_all = [torch.empty_like(ft_loc) for _ in range(world)]
        dist.all_gather(ft_all, ft_loc)
        fi_loc = eng.vision_forward(images[rank * B:(rank + 1) * B].contiguous(), P_v)
    if rank == 0:
        ref = _build_big(B * world, C).to(dev)  # fresh engine: unsharded reference
        ref.shard_classes = False
        with torch.no_grad():
            reng = ref._clip_ref[0].engine()
            ref._register_classes(dev)
            rP_v, rP_t = ref.prompt_stacks()
            out["text_feat_max_abs"] = float((torch.cat(ft_all) - reng.text_forward(rP_t, True)).abs().max())
            out["img_feat_max_abs"] = float((fi_loc - reng.vision_forward(images, rP_v)[:B]).abs().max())
        ref.zero_grad(set_to_none=True)
        loss_f, logits_f = ref.forward_backward(images, labels)
        out["loss_fresh_single"] = float(loss_f)
        out["logits_vs_fresh_max_abs"] = float((torch.cat(gathered) - logits_f).abs().max())
        model.shard_classes = False
        model._clip_ref[0].engine().class_key = None
        model.zero_grad(set_to_none=True)
        loss1, logits1 = model.forward_backward(images, labels)
        torch.cuda.synchronize()
        out["logits_reused_vs_fresh_max_abs"] = float((logits1 - logits_f).abs().max())
        out["loss_sharded"], out["loss_single"] = float(loss), float(loss1)
        out["logits_max_abs"] = float((torch.cat(gathered) - logits1).abs().max())
        worst = 1.0
        for n, p in model.named_parameters():
            if p.requires_grad:
                a, b = grads[n].flatten().double(), p.grad.flatten().double()
                cos = float((a @ b) / (a.norm() * b.norm()))
                rel = float((a - b).norm() / b.norm())
                out["grad/" + n] = {"cos": cos, "rel": rel}
                worst = min(worst, cos)
        out["ok"] = bool(abs(out["loss_sharded"] - out["loss_single"]) < 1e-3 and out["logits_max_abs"] < 2e-2 and worst > 0.9999)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"dist_check_w{world}.json"), "w"), indent=1)
        print(json.dumps({k: v for k, v in out.items() if not k.startswith("grad/")}), "worst grad cos", worst, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not out["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
