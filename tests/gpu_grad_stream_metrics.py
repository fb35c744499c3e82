"""Effect of the bf16 gradient stream (option grad_stream_bf16) on the prompt gradients: per-tensor cosine / rel-L2 against the
REFERENCE golden vectors (tests/golden, made by oracle/make_golden.py from the unmodified reference) with the stream in
fp32 (0) and bf16 (1), and the two native results against each other.  Run on the B200 box:
    python tests/gpu_grad_stream_metrics.py > gpurun_out/r02_grad_stream_bf16.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mudpt_b200 import _lib
from oracle import mudpt_oracle as orc  # the checker
from tests import golden_util as gu


def run(model, eng, c, opt):
    _lib.check(eng.lib.mudpt_set_option(eng.h, b"grad_stream_bf16", opt), eng.h)
    model.zero_grad(set_to_none=True)
    loss, logits = model.forward_backward(c["image"].cuda(), c["labels"].cuda())
    torch.cuda.synchronize()
    params = dict(model.named_parameters())
    return float(loss), logits.detach().cpu(), {k: params[k].grad.detach().cpu().clone() for k in orc.TRAINABLE}


print("# prompt gradients against the reference golden: gradient stream fp32 (grad_stream_bf16 = 0) vs bf16 (= 1)")
print("# case  tensor  cos(fp32 stream)  rel-L2(fp32 stream)  cos(bf16 stream)  rel-L2(bf16 stream)  rel-L2(bf16 vs fp32 stream)")
for name in ("tiny_a", "tiny_b", "tiny_c", "tiny_d", "vitb16_cfg1"):
    try:
        c = gu.load(name)
    except Exception as e:  # noqa: BLE001
        print(f"# {name}: not available ({e})")
        continue
    model, _ = gu.build_model(c, "cuda")
    eng = model._clip_ref[0].engine()
    g = c["golden"]
    r0 = run(model, eng, c, 0)
    r1 = run(model, eng, c, 1)
    print(f"{name}: loss {r0[0]:.6f} / {r1[0]:.6f} (golden {float(g['loss']):.6f}); logits identical: {bool(torch.equal(r0[1], r1[1]))}")
    worst = [1.0, 0.0, 1.0, 0.0, 0.0]
    for k in orc.TRAINABLE:
        ref = torch.from_numpy(g["grad/" + k])
        if not ref.numel() or float(ref.norm()) == 0:
            continue
        m0, m1 = orc.metrics(r0[2][k], ref), orc.metrics(r1[2][k], ref)
        d = float((r1[2][k] - r0[2][k]).norm() / (r0[2][k].norm() + 1e-30))
        print(f"  {k:55s} {m0['cos']:.6f} {m0['rel_l2']:.4f}   {m1['cos']:.6f} {m1['rel_l2']:.4f}   {d:.4f}")
        worst = [min(worst[0], m0["cos"]), max(worst[1], m0["rel_l2"]), min(worst[2], m1["cos"]), max(worst[3], m1["rel_l2"]), max(worst[4], d)]
    print(f"  worst: cos {worst[0]:.6f} rel {worst[1]:.4f} (fp32 stream)   cos {worst[2]:.6f} rel {worst[3]:.4f} (bf16 stream)   "
          f"bf16 vs fp32 stream rel {worst[4]:.4f}")
