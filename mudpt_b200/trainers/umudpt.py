"""UMuDPT / UUMuDPT on the native towers -- drop-ins for the reference's trainers/umudpt.py and
trainers/uumudpt.py.

The towers are the MuDPT towers (the UMuDPT / UUMuDPT blocks, clip/model.py:302-400, splice exactly
like the MuDPT block); what differs is where the spliced prompts come from: a one-block
`LightTransformer` mixes [ctx; deep_prompts] into the visual prompts (trainers/umudpt.py:172-178),
and in UUMuDPT the vision side owns prompts of its own whose LightTransformer output is added to the
text deep prompts (clip/model.py:600-664, trainers/uumudpt.py:223-224).  That is [n_ctx x depth]-token
algebra: it stays in torch autograd and feeds the same two [depth, n_ctx, width] prompt stacks the C ABI
takes, so the fused step, the class-sharded multi-GPU path and the cached-text-feature inference of
`mudpt_b200.trainers.mudpt.CustomCLIP` are inherited unchanged.

  UMuDPTPromptLearner(cfg, classnames, clip_model) ... trainers/umudpt.py:82-180
  CustomCLIP(cfg, classnames, clip_model) ............ trainers/umudpt.py:207-229
  UMuDPT (trainer) ................................... trainers/umudpt.py:232-346
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import clip
from ..clip.model import LightTransformer, _stack_with_ln_pre
from . import mudpt as _m
from .mudpt import TRAINER_REGISTRY, TextEncoder, build_lr_scheduler, build_optimizer, load_pretrained_weights


def load_clip_to_cpu(cfg):
    """trainers/umudpt.py:24-41 (random-init when no BACKBONE.PATH is given: no checkpoint exists on the boxes)."""
    return _m.load_clip_to_cpu(cfg)


class UMuDPTPromptLearner(nn.Module):
    CFG_NODE = "UMUDPT"

    def __init__(self, cfg, classnames, clip_model, tokenizer=None):
        super().__init__()
        tokenize = tokenizer if tokenizer is not None else clip.tokenize
        node = getattr(cfg.TRAINER, self.CFG_NODE)
        n_cls = len(classnames)
        n_ctx = node.N_CTX
        ctx_init = node.CTX_INIT
        dtype = clip_model.dtype
        ctx_dim = clip_model.ln_final.weight.shape[0]
        clip_imsize = clip_model.visual.input_resolution
        cfg_imsize = cfg.INPUT.SIZE[0]
        visual_ctx_dim = clip_model.visual.positional_embedding.shape[1]

        assert node.DEEP_PROMPT_DEPTH > 0, "PROMPT_DEPTH should be > 0"
        self.deep_prompts_depth = node.DEEP_PROMPT_DEPTH
        assert cfg_imsize == clip_imsize, f"cfg_imsize ({cfg_imsize}) must equal to clip_imsize ({clip_imsize})"

        if ctx_init:
            ctx_init = ctx_init.replace("_", " ")
            prompt = tokenize(ctx_init)
            with torch.no_grad():
                embedding = clip_model.token_embedding(prompt.to(clip_model.token_embedding.weight.device).long()).type(dtype)
            ctx_vectors = embedding[0, 1: 1 + n_ctx, :].clone()
            prompt_prefix = " ".join(ctx_init.split()[:n_ctx])
        else:
            ctx_vectors = torch.empty(n_ctx, ctx_dim, dtype=dtype)
            nn.init.normal_(ctx_vectors, std=0.02)
            prompt_prefix = " ".join(["X"] * n_ctx)
        self.ctx = nn.Parameter(ctx_vectors)

        self.deep_prompts = nn.Parameter(torch.empty(self.deep_prompts_depth - 1, n_ctx, ctx_dim))
        nn.init.normal_(self.deep_prompts, std=0.02)

        # light transformer for t2v prompts
        self.ln_pre = nn.LayerNorm(ctx_dim)
        self.self_attn = LightTransformer(d_model=ctx_dim, n_head=ctx_dim // 64)
        self.ln_post = nn.LayerNorm(ctx_dim)
        self.visual_proj = nn.Linear(in_features=ctx_dim, out_features=visual_ctx_dim)

        classnames = [name.replace("_", " ") for name in classnames]
        prompts = [prompt_prefix + " " + name + "." for name in classnames]
        tokenized_prompts = torch.cat([tokenize(p) for p in prompts], dim=0)
        with torch.no_grad():
            embedding = clip_model.token_embedding(
                tokenized_prompts.to(clip_model.token_embedding.weight.device).long()).type(dtype)
        self.register_buffer("token_prefix", embedding[:, :1, :].clone())          # SOS
        self.register_buffer("token_suffix", embedding[:, 1 + n_ctx:, :].clone())  # CLS . EOS ~

        self.n_cls = n_cls
        self.n_ctx = n_ctx
        self.dtype = dtype
        self.ctx_dim = ctx_dim
        self.tokenized_prompts = tokenized_prompts

    def construct_prompts(self, ctx, prefix, suffix, label=None):
        if label is not None:
            prefix = prefix[label]
            suffix = suffix[label]
        return torch.cat([prefix, ctx, suffix], dim=1)

    def visual_prompts(self):
        """[depth, n_ctx, visual width] (trainers/umudpt.py:172-178)."""
        v = torch.cat([self.ctx.unsqueeze(0), self.deep_prompts], dim=0)
        v = self.ln_pre(v)
        v = self.self_attn(v.permute(1, 0, 2)).permute(1, 0, 2)
        return self.visual_proj(self.ln_post(v))

    def forward(self):
        """Reference API (trainers/umudpt.py:162-180): materialises the class prompts.  The fused path does not call this."""
        ctx = self.ctx
        if ctx.dim() == 2:
            ctx = ctx.unsqueeze(0).expand(self.n_cls, self.n_ctx, self.ctx_dim)
        ctx = self.construct_prompts(ctx, self.token_prefix, self.token_suffix)
        return ctx, self.deep_prompts, self.visual_prompts()


class CustomCLIP(_m.CustomCLIP):
    """Same wiring as the MuDPT CustomCLIP (fused step, sharding, cached inference); only the prompt stacks differ."""

    LEARNER_ATTR = "umudpt_prompt_learner"
    LEARNER_CLS = UMuDPTPromptLearner

    def __init__(self, cfg, classnames, clip_model, tokenizer=None):
        nn.Module.__init__(self)
        setattr(self, self.LEARNER_ATTR, self.LEARNER_CLS(cfg, classnames, clip_model, tokenizer=tokenizer))
        self.tokenized_prompts = self.mudpt_prompt_learner.tokenized_prompts
        self.text_encoder = TextEncoder(clip_model)
        self.image_encoder = clip_model.visual
        self.logit_scale = clip_model.logit_scale
        self.dtype = clip_model.dtype
        self.deep_prompts_depth = self.mudpt_prompt_learner.deep_prompts_depth
        object.__setattr__(self, "_clip_ref", [clip_model])
        self.truncate_text_to_eot = os.environ.get("MUDPT_TEXT_FULL_LENGTH", "0") != "1"
        self.shard_classes = True
        self._cached_text_features = None
        self._replicas_synced = False
        self.overlap_towers = os.environ.get("MUDPT_OVERLAP_TOWERS", "1") != "0"
        self._loss_host, self._loss_event, self._loss_pending = None, None, False

    @property
    def mudpt_prompt_learner(self):  # the inherited code paths address the learner by this name
        return getattr(self, self.LEARNER_ATTR)

    def prompt_stacks(self):
        pl, ve = self.mudpt_prompt_learner, self.image_encoder
        vp = pl.visual_prompts()
        P_v = _stack_with_ln_pre(ve, vp[:1], vp[1:])
        pos = self.text_encoder.positional_embedding[1:1 + pl.n_ctx]
        P_t = torch.cat([(pl.ctx + pos).unsqueeze(0), pl.deep_prompts], dim=0).float()
        return P_v, P_t


def _freeze(model, keep_visual_ctx: bool):
    for name, param in model.named_parameters():
        if "prompt_learner" not in name:
            param.requires_grad_(keep_visual_ctx and "visual_ctx" in name)


@TRAINER_REGISTRY.register()
class UMuDPT(_m.MuDPT):
    MODEL_NAME = "UnifiedMultimodalDeepPromptTuning"
    CFG_NODE = "UMUDPT"
    KEEP_VISUAL_CTX = False
    CUSTOM_CLIP = CustomCLIP

    def check_cfg(self, cfg):
        assert getattr(cfg.TRAINER, self.CFG_NODE).PREC in ["fp16", "fp32", "amp"]

    def build_model(self):
        cfg = self.cfg
        classnames = self.dm.dataset.classnames if hasattr(self, "dm") else self._classnames
        print(f"Loading CLIP (backbone: {cfg.MODEL.BACKBONE.NAME})")
        clip_model = load_clip_to_cpu(cfg)
        clip_model.float()
        print("Building custom CLIP")
        self.model = self.CUSTOM_CLIP(cfg, classnames, clip_model)
        print("Turning off gradients in both the image and the text encoder")
        _freeze(self.model, self.KEEP_VISUAL_CTX)
        enabled = {name for name, p in self.model.named_parameters() if p.requires_grad}
        print(f"Parameters to be updated: {enabled}")
        if getattr(cfg.MODEL, "INIT_WEIGHTS", ""):
            load_pretrained_weights(self.model.mudpt_prompt_learner, cfg.MODEL.INIT_WEIGHTS)
        self.model.to(self.device)
        self.optim = build_optimizer(self.model, cfg.OPTIM)
        self.sched = build_lr_scheduler(self.optim, cfg.OPTIM)
        self.register_model(self.MODEL_NAME, self.model, self.optim, self.sched)
        self.scaler = None
