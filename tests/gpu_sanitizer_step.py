"""One process for `compute-sanitizer --tool memcheck` on the B200 box: the row kernels in every form (LayerNorm forward, both
LayerNorm backward kernels incl. the bulk-copy pipeline, splice, im2col) and one full train step at BASELINE config 1 (ViT-B/16,
4 images, 100 classes: every main-loop kernel at its real widths), checked against the reference golden as usual.

    compute-sanitizer --tool memcheck --error-exitcode 3 python tests/gpu_sanitizer_step.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import mudpt_oracle as orc  # the checker
from tests import golden_util as gu
from tests import gpu_bringup as bring

t0 = time.time()
res = {}
bring.group_rowops(res)
torch.cuda.synchronize()
print(f"row kernels done ({time.time() - t0:.0f} s)", flush=True)
c = gu.load("vitb16_cfg1")
model, _ = gu.build_model(c, "cuda")
loss, logits = model.forward_backward(c["image"].cuda(), c["labels"].cuda())
torch.cuda.synchronize()
g = c["golden"]
err = float((logits.cpu() - torch.from_numpy(g["logits"])).abs().max())
worst = 1.0
params = dict(model.named_parameters())
for k in orc.TRAINABLE:
    ref = torch.from_numpy(g["grad/" + k])
    if ref.numel() and float(ref.norm()) > 0:
        worst = min(worst, orc.metrics(params[k].grad.cpu(), ref)["cos"])
eng = model._clip_ref[0].engine()
print(f"config-1 step done ({time.time() - t0:.0f} s): loss {float(loss):.5f} (golden {float(g['loss']):.5f}), logit max-abs err {err:.4f}, "
      f"min prompt-gradient cosine {worst:.6f}, {eng.launch_count()} native kernel launches", flush=True)
assert err <= 0.05 and worst >= 0.999
