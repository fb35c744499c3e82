"""CoCoOp on the shared native CLIP kernels -- drop-in for the reference's trainers/cocoop.py
(BASELINE config 4: instance-conditioned prompts, B x C text sequences per step).

Same public classes, constructor signatures, parameter / buffer names and return values:

  PromptLearner(cfg, classnames, clip_model) ... trainers/cocoop.py:66-163
  TextEncoder(clip_model) ...................... trainers/cocoop.py:43-64
  CustomCLIP(cfg, classnames, clip_model) ...... trainers/cocoop.py:166-198
  CoCoOp (trainer) ............................. trainers/cocoop.py:201-320

What changed underneath: the plain (cfg=None) CLIP towers run in libmudpt_b200.so; the reference's
Python loop over the images of a batch (one text-tower pass of C sequences per image, :187-192)
is ONE native text-tower call over all B x C sequences, truncated to max(eot)+1 tokens (exact under
the causal mask); the vision tower is forward-only (it holds nothing trainable).  The meta-net, the
prompt assembly and the per-image cosine logits (B x C x e multiply-adds) stay in torch autograd,
which delivers the gradients of the 5 trainable tensors from the dense d(prompts) the native text
backward returns.
"""
from __future__ import annotations

import os.path as osp
from collections import OrderedDict

import torch
import torch.nn as nn
from torch.nn import functional as F

from .. import clip
from ..engine import TextTowerDenseFn
from .mudpt import TRAINER_REGISTRY, TrainerX, build_lr_scheduler, build_optimizer, load_checkpoint, load_pretrained_weights


def load_clip_to_cpu(cfg=None):
    """trainers/cocoop.py:22-40: a plain CLIP (no cfg -> plain blocks).  Without BACKBONE.PATH the model
    is random-initialised (no checkpoint exists on the build / GPU boxes)."""
    path = getattr(cfg.MODEL.BACKBONE, "PATH", "")
    if path:
        try:
            sd = torch.jit.load(path, map_location="cpu").state_dict()
        except RuntimeError:
            sd = torch.load(path, map_location="cpu")
        return clip.build_model(sd)
    from ..synthetic import ARCHS
    return clip.CLIP(*ARCHS[cfg.MODEL.BACKBONE.NAME].astuple(), None).float().eval()


class TextEncoder(nn.Module):
    def __init__(self, clip_model):
        super().__init__()
        self.transformer = clip_model.transformer
        self.positional_embedding = clip_model.positional_embedding
        self.ln_final = clip_model.ln_final
        self.text_projection = clip_model.text_projection
        self.dtype = clip_model.dtype
        object.__setattr__(self, "_clip_ref", [clip_model])
        self.truncate_to_eot = True

    def forward(self, prompts, tokenized_prompts):
        """prompts [N, 77, d] (N = C, or B*C batched), tokenized_prompts [N, 77] -> text features [N, e]."""
        engine = self._clip_ref[0].engine(prompts.device)
        eot = tokenized_prompts.argmax(dim=-1).cpu()
        seq_len = int(eot.max()) + 1 if self.truncate_to_eot else prompts.shape[1]
        deep = torch.zeros(0, 0, prompts.shape[-1], device=prompts.device)
        return TextTowerDenseFn.apply(engine, prompts.float(), eot, deep, seq_len)


class PromptLearner(nn.Module):
    def __init__(self, cfg, classnames, clip_model, tokenizer=None, name_len=None):
        super().__init__()
        tokenize = tokenizer if tokenizer is not None else clip.tokenize
        n_cls = len(classnames)
        n_ctx = cfg.TRAINER.COCOOP.N_CTX
        ctx_init = cfg.TRAINER.COCOOP.CTX_INIT
        dtype = clip_model.dtype
        ctx_dim = clip_model.ln_final.weight.shape[0]
        vis_dim = clip_model.visual.output_dim
        clip_imsize = clip_model.visual.input_resolution
        cfg_imsize = cfg.INPUT.SIZE[0]
        assert cfg_imsize == clip_imsize, f"cfg_imsize ({cfg_imsize}) must equal to clip_imsize ({clip_imsize})"

        if ctx_init:
            ctx_init = ctx_init.replace("_", " ")
            n_ctx = len(ctx_init.split(" "))
            prompt = tokenize(ctx_init)
            with torch.no_grad():
                embedding = clip_model.token_embedding(prompt.to(clip_model.token_embedding.weight.device).long()).type(dtype)
            ctx_vectors = embedding[0, 1: 1 + n_ctx, :].clone()
            prompt_prefix = ctx_init
        else:
            ctx_vectors = torch.empty(n_ctx, ctx_dim, dtype=dtype)
            nn.init.normal_(ctx_vectors, std=0.02)
            prompt_prefix = " ".join(["X"] * n_ctx)
        self.ctx = nn.Parameter(ctx_vectors)

        self.meta_net = nn.Sequential(OrderedDict([
            ("linear1", nn.Linear(vis_dim, vis_dim // 16)),
            ("relu", nn.ReLU(inplace=True)),
            ("linear2", nn.Linear(vis_dim // 16, ctx_dim)),
        ]))

        classnames = [name.replace("_", " ") for name in classnames]
        prompts = [prompt_prefix + " " + name + "." for name in classnames]
        tokenized_prompts = torch.cat([tokenize(p) for p in prompts])
        with torch.no_grad():
            embedding = clip_model.token_embedding(
                tokenized_prompts.to(clip_model.token_embedding.weight.device).long()).type(dtype)
        self.register_buffer("token_prefix", embedding[:, :1, :].clone())          # SOS
        self.register_buffer("token_suffix", embedding[:, 1 + n_ctx:, :].clone())  # CLS, EOS

        self.n_cls = n_cls
        self.n_ctx = n_ctx
        self.tokenized_prompts = tokenized_prompts
        self.name_lens = name_len

    def construct_prompts(self, ctx, prefix, suffix, label=None):
        if label is not None:
            prefix = prefix[label]
            suffix = suffix[label]
        return torch.cat([prefix, ctx, suffix], dim=1)

    def forward(self, im_features):
        """[B, e] -> [B, n_cls, 77, d] (trainers/cocoop.py:149-163), built with one broadcasted cat."""
        bias = self.meta_net(im_features).unsqueeze(1)          # (batch, 1, ctx_dim)
        ctx_shifted = self.ctx.unsqueeze(0) + bias              # (batch, n_ctx, ctx_dim)
        B = ctx_shifted.shape[0]
        prefix = self.token_prefix.unsqueeze(0).expand(B, -1, -1, -1)
        suffix = self.token_suffix.unsqueeze(0).expand(B, -1, -1, -1)
        ctx = ctx_shifted.unsqueeze(1).expand(-1, self.n_cls, -1, -1)
        return torch.cat([prefix, ctx, suffix], dim=2)


class CustomCLIP(nn.Module):
    def __init__(self, cfg, classnames, clip_model, tokenizer=None):
        super().__init__()
        self.prompt_learner = PromptLearner(cfg, classnames, clip_model, tokenizer=tokenizer)
        self.tokenized_prompts = self.prompt_learner.tokenized_prompts
        self.image_encoder = clip_model.visual
        self.text_encoder = TextEncoder(clip_model)
        self.logit_scale = clip_model.logit_scale
        self.dtype = clip_model.dtype
        object.__setattr__(self, "_clip_ref", [clip_model])

    def forward(self, image, label=None):
        tokenized_prompts = self.tokenized_prompts
        logit_scale = self.logit_scale.exp()

        image_features = self.image_encoder(image.type(self.dtype))
        image_features = image_features / image_features.norm(dim=-1, keepdim=True)

        prompts = self.prompt_learner(image_features)                       # [B, C, 77, d]
        B, C = prompts.shape[0], prompts.shape[1]
        # all B x C sequences in one native text-tower pass (the reference loops over the images)
        text_features = self.text_encoder(prompts.reshape(B * C, prompts.shape[2], prompts.shape[3]),
                                          tokenized_prompts.repeat(B, 1))
        text_features = text_features.view(B, C, -1)
        text_features = text_features / text_features.norm(dim=-1, keepdim=True)
        # per-image cosine logits: an fp32 multiply + row sum (no library GEMM, so no TF32 / algorithm dependence)
        logits = logit_scale * (image_features.unsqueeze(1) * text_features).sum(dim=-1)

        if self.prompt_learner.training and label is not None:
            return F.cross_entropy(logits, label)
        return logits


@TRAINER_REGISTRY.register()
class CoCoOp(TrainerX):
    def check_cfg(self, cfg):
        assert cfg.TRAINER.COCOOP.PREC in ["fp16", "fp32", "amp"]

    def build_model(self):
        cfg = self.cfg
        classnames = self.dm.dataset.classnames if hasattr(self, "dm") else self._classnames
        print(f"Loading CLIP (backbone: {cfg.MODEL.BACKBONE.NAME})")
        clip_model = load_clip_to_cpu(cfg)
        clip_model.float()

        print("Building custom CLIP")
        self.model = CustomCLIP(cfg, classnames, clip_model)

        print("Turning off gradients in both the image and the text encoder")
        name_to_update = "prompt_learner"
        for name, param in self.model.named_parameters():
            if name_to_update not in name:
                param.requires_grad_(False)
        enabled = {name for name, p in self.model.named_parameters() if p.requires_grad}
        print(f"Parameters to be updated: {enabled}")

        if getattr(cfg.MODEL, "INIT_WEIGHTS", ""):
            load_pretrained_weights(self.model.prompt_learner, cfg.MODEL.INIT_WEIGHTS)

        self.model.to(self.device)
        # NOTE: only give prompt_learner to the optimizer
        self.optim = build_optimizer(self.model.prompt_learner, cfg.OPTIM)
        self.sched = build_lr_scheduler(self.optim, cfg.OPTIM)
        self.register_model("prompt_learner", self.model.prompt_learner, self.optim, self.sched)
        self.scaler = None  # "amp" needs no loss scaling here: bf16 operands, fp32 accumulation and master state

    def forward_backward(self, batch):
        image, label = self.parse_batch_train(batch)
        loss = self.model(image, label)
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        loss_summary = {"loss": loss.item()}
        if (self.batch_idx + 1) == self.num_batches:
            self.update_lr()
        return loss_summary

    def parse_batch_train(self, batch):
        input = batch["img"].to(self.device)
        label = batch["label"].to(self.device)
        return input, label

    def load_model(self, directory, epoch=None):
        if not directory:
            print("Note that load_model() is skipped as no Pretrained model is given")
            return
        names = self.get_model_names()
        model_file = "model-best.pth.tar"
        if epoch is not None:
            model_file = "model.pth.tar-" + str(epoch)
        for name in names:
            model_path = osp.join(directory, name, model_file)
            if not osp.exists(model_path):
                raise FileNotFoundError('Model not found at "{}"'.format(model_path))
            checkpoint = load_checkpoint(model_path)
            state_dict = checkpoint["state_dict"]
            epoch = checkpoint["epoch"]
            # ignore the fixed token vectors: they are recomputed from the current class names
            state_dict.pop("token_prefix", None)
            state_dict.pop("token_suffix", None)
            print('Loading weights to {} from "{}" (epoch = {})'.format(name, model_path, epoch))
            self._models[name].load_state_dict(state_dict, strict=False)
