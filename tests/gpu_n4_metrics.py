"""Actual N4 (UMuDPT / UUMuDPT) errors against the reference golden, to size the tolerances of the GPU test."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import golden_util as gu
from oracle import mudpt_oracle as orc
for name in ("umudpt_tiny", "uumudpt_tiny"):
    c = gu.load_variant(name); g = c["golden"]
    model, _ = gu.build_variant_model(c, "cuda")
    image, labels = c["image"].cuda(), c["labels"].cuda()
    model.zero_grad(set_to_none=True)
    loss, logits = model.forward_backward(image, labels)
    torch.cuda.synchronize()
    print(name, "dloss", abs(float(loss) - float(g["loss"])), "logits max_abs", float((logits.detach().cpu() - torch.from_numpy(g["logits"])).abs().max()),
          "logit scale", float(torch.from_numpy(g["logits"]).abs().max()))
    worst = (1.0, None)
    for n, p in model.named_parameters():
        if p.requires_grad:
            m = orc.metrics(p.grad.cpu(), torch.from_numpy(g["grad/" + n]))
            if m["cos"] < worst[0]: worst = (m["cos"], (n, m))
    print("  worst grad", worst)
