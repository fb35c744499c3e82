"""TEST INFRASTRUCTURE: a CPU stand-in for mudpt_b200.engine.Engine built on the oracle, so that
the host-side logic (prompt stacks, autograd plumbing, class sharding, collectives) can be
tested without a GPU (`-m "not gpu"`, gloo world_size 2).  Never imported by the product."""
import torch

from oracle import mudpt_oracle as orc


class OracleEngine:
    def __init__(self, sd, arch, n_ctx, depth, device=torch.device("cpu")):
        self.sd = {k: v.detach() for k, v in sd.items()}
        self.arch = dict(zip(["embed_dim", "image_resolution", "vision_layers", "vision_width", "vision_patch_size",
                              "context_length", "vocab_size", "transformer_width", "transformer_heads",
                              "transformer_layers"], arch.astuple()))
        self.n_ctx, self.depth, self.device = n_ctx, depth, device
        self.vision_gen = self.text_gen = 0
        self.n_classes = self.text_len = 0
        self.class_key = None
        self._v = self._t = None

    def vision_forward(self, images, prompts):
        P = prompts.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            f = orc.vision_features_from_stack(self.sd, images.float(), P)
        self._v = (P, f)
        self.vision_gen += 1
        return f.detach()

    def vision_backward(self, d):
        P, f = self._v
        return torch.autograd.grad(f, P, d)[0]

    def text_set_classes(self, embeddings, eot, seq_len):
        self._emb, self._eot = embeddings.detach().float(), eot.long()
        self.n_classes, self.text_len = embeddings.shape[0], int(seq_len)
        self.class_key = None

    def text_forward(self, prompts, splice_layer0=True):
        P = prompts.detach().clone().requires_grad_(True)
        x0 = self._emb.clone().requires_grad_(True)
        with torch.enable_grad():
            f = orc.text_features_from_stack(self.sd, x0, self._eot, P, self.text_len, 0 if splice_layer0 else 1)
        self._t = (P, x0, f)
        self.text_gen += 1
        return f.detach()

    def text_backward(self, d, want_dx0=False):
        P, x0, f = self._t
        gP, gx = torch.autograd.grad(f, [P, x0], d, allow_unused=True)
        if gP is None:
            gP = torch.zeros_like(P)
        return gP, (gx[:, :self.text_len].contiguous() if want_dx0 else None)

    def logits_head(self, f_img, f_txt, labels, inv_global_batch, want_grads):
        fi = f_img.detach().clone().requires_grad_(True)
        ft = f_txt.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            logits, loss = orc.logits_and_loss(fi, ft, self.sd["logit_scale"], labels, inv_global_batch)
        if labels is None:
            return logits.detach(), torch.zeros(()), None, None
        di, dt = torch.autograd.grad(loss, [fi, ft]) if want_grads else (None, None)
        return logits.detach(), loss.detach(), di, dt

    def logits_backward(self, f_img, f_txt, dlogits):
        fi = f_img.detach().clone().requires_grad_(True)
        ft = f_txt.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            logits, _ = orc.logits_and_loss(fi, ft, self.sd["logit_scale"])
        return torch.autograd.grad(logits, [fi, ft], dlogits)

    def launch_count(self):
        return 0


def attach(model, case):
    """Route a mudpt_b200 CustomCLIP (built on CPU) through the oracle-backed engine."""
    eng = OracleEngine(case["sd"], case["arch"], case["n_ctx"], case["depth"])
    model._clip_ref[0].engine = lambda device=None: eng
    return eng
