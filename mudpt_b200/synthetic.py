"""Deterministic synthetic weights / tokens / inputs for the MuDPT hot path.

There are no pretrained CLIP checkpoints and no datasets on the build or GPU boxes
(SURVEY.md section 8c/8d), so every run uses random-init weights of the named architecture and
synthetic inputs.  Everything here is a pure function of integer seeds and uses numpy's
PCG64 generator (scalar C code, bit-stable across hosts of the same image), so the golden
fixtures made in the build container (oracle/make_golden.py, from the *reference*) can be
reproduced bit-for-bit on the GPU box without shipping 500 MB of weights.

The state-dict keys are the reference's `CustomCLIP.state_dict()` names
(trainers/mudpt.py:159-168, clip/model.py:499-524, 667-779); the distributions follow the
reference constructors (`CLIP.initialize_parameters`, clip/model.py:781-808; torch default
inits for the vision tower; N(0, 0.02^2) prompts, clip/model.py:513,517 and
trainers/mudpt.py:69,79) except that LayerNorm affine parameters and biases are perturbed
away from 1/0 so that parity tests are sensitive to them.
"""
from __future__ import annotations

import math
import re
import zlib
from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np
import torch

SOT_TOKEN = 49406  # clip/simple_tokenizer.py:73-74
EOT_TOKEN = 49407
CONTEXT_LENGTH = 77


@dataclass(frozen=True)
class Arch:
    embed_dim: int
    image_resolution: int
    vision_layers: int
    vision_width: int
    vision_patch_size: int
    context_length: int
    vocab_size: int
    transformer_width: int
    transformer_heads: int
    transformer_layers: int

    @property
    def vision_heads(self) -> int:
        return self.vision_width // 64  # clip/model.py:696

    @property
    def n_patches(self) -> int:
        return (self.image_resolution // self.vision_patch_size) ** 2

    def astuple(self):
        return (self.embed_dim, self.image_resolution, self.vision_layers, self.vision_width,
                self.vision_patch_size, self.context_length, self.vocab_size,
                self.transformer_width, self.transformer_heads, self.transformer_layers)


ARCHS: Dict[str, Arch] = {
    "ViT-B/16": Arch(512, 224, 12, 768, 16, 77, 49408, 512, 8, 12),
    "ViT-B/32": Arch(512, 224, 12, 768, 32, 77, 49408, 512, 8, 12),
    "ViT-L/14": Arch(768, 224, 24, 1024, 14, 77, 49408, 768, 12, 12),
    # small shapes for fast tests: same code path, heads of 64
    "tiny": Arch(64, 32, 3, 128, 16, 77, 49408, 64, 1, 3),
    "tiny2": Arch(128, 48, 2, 192, 16, 77, 49408, 128, 2, 2),
}


# ------------------------------------------------------------------------------------------
# tokens
# ------------------------------------------------------------------------------------------

_PIECE = re.compile(r"[a-z]+|[0-9]|[^\sa-z0-9]")


def synthetic_tokenize(texts, context_length: int = CONTEXT_LENGTH) -> torch.Tensor:
    """Stand-in for clip.tokenize (clip/clip.py:199-239) when no BPE vocabulary is on the box.

    Word-level: lower-cased alphabetic runs, single digits and single punctuation marks each
    become one token id in [1000, 41000) (crc32 of the piece); SOT/EOT ids and the zero padding
    are the reference's, so `argmax` still finds the EOT position (trainers/mudpt.py:154).
    The BPE tokenizer itself is out of scope: its output is an
    *input* of the hot path.  Pass the reference's `clip.tokenize` to the prompt learner to
    get real token ids.
    """
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.int32)
    for i, t in enumerate(texts):
        ids = [SOT_TOKEN] + [1000 + zlib.crc32(p.encode()) % 40000 for p in _PIECE.findall(t.lower())] + [EOT_TOKEN]
        if len(ids) > context_length:
            raise RuntimeError(f"Input {t} is too long for context length {context_length}")
        out[i, :len(ids)] = torch.tensor(ids, dtype=torch.int32)
    return out


# ------------------------------------------------------------------------------------------
# weights
# ------------------------------------------------------------------------------------------

class _Rng:
    def __init__(self, seed: int):
        self.g = np.random.default_rng(seed)

    def normal(self, shape, std=1.0):
        return torch.from_numpy((self.g.standard_normal(shape, dtype=np.float32) * np.float32(std)))

    def uniform(self, shape, bound):
        return torch.from_numpy(((self.g.random(shape, dtype=np.float32) * 2 - 1) * np.float32(bound)))


def _block(sd, pfx, d, r: _Rng, text: bool, layers: int):
    if text:  # clip/model.py:795-803
        sd[pfx + "attn.in_proj_weight"] = r.normal((3 * d, d), d ** -0.5)
        sd[pfx + "attn.out_proj.weight"] = r.normal((d, d), (d ** -0.5) * ((2 * layers) ** -0.5))
        sd[pfx + "mlp.c_fc.weight"] = r.normal((4 * d, d), (2 * d) ** -0.5)
        sd[pfx + "mlp.c_proj.weight"] = r.normal((d, 4 * d), (d ** -0.5) * ((2 * layers) ** -0.5))
    else:  # torch defaults: xavier_uniform in-proj, kaiming_uniform(a=sqrt(5)) Linear
        sd[pfx + "attn.in_proj_weight"] = r.uniform((3 * d, d), math.sqrt(6.0 / (4 * d)))
        sd[pfx + "attn.out_proj.weight"] = r.uniform((d, d), d ** -0.5)
        sd[pfx + "mlp.c_fc.weight"] = r.uniform((4 * d, d), d ** -0.5)
        sd[pfx + "mlp.c_proj.weight"] = r.uniform((d, 4 * d), (4 * d) ** -0.5)
    sd[pfx + "attn.in_proj_bias"] = r.normal((3 * d,), 0.02)
    sd[pfx + "attn.out_proj.bias"] = r.normal((d,), 0.02)
    sd[pfx + "mlp.c_fc.bias"] = r.uniform((4 * d,), d ** -0.5)
    sd[pfx + "mlp.c_proj.bias"] = r.uniform((d,), (4 * d) ** -0.5)
    for ln in ("ln_1", "ln_2"):
        sd[pfx + ln + ".weight"] = 1.0 + r.normal((d,), 0.1)
        sd[pfx + ln + ".bias"] = r.normal((d,), 0.1)


def synthetic_clip_state_dict(arch: Arch, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Frozen CLIP weights under the reference `CLIP.state_dict()` names (fp32, CPU)."""
    r = _Rng(seed)
    a = arch
    sd: Dict[str, torch.Tensor] = {}
    dv, dt, e = a.vision_width, a.transformer_width, a.embed_dim
    p = a.vision_patch_size
    V = "visual."
    sd[V + "conv1.weight"] = r.uniform((dv, 3, p, p), (3 * p * p) ** -0.5)
    sd[V + "class_embedding"] = r.normal((dv,), dv ** -0.5)
    sd[V + "positional_embedding"] = r.normal((a.n_patches + 1, dv), dv ** -0.5)
    sd[V + "ln_pre.weight"] = 1.0 + r.normal((dv,), 0.1)
    sd[V + "ln_pre.bias"] = r.normal((dv,), 0.1)
    for i in range(a.vision_layers):
        _block(sd, f"{V}transformer.resblocks.{i}.", dv, r, False, a.vision_layers)
    sd[V + "ln_post.weight"] = 1.0 + r.normal((dv,), 0.1)
    sd[V + "ln_post.bias"] = r.normal((dv,), 0.1)
    sd[V + "proj"] = r.normal((dv, e), dv ** -0.5)
    for i in range(a.transformer_layers):
        _block(sd, f"transformer.resblocks.{i}.", dt, r, True, a.transformer_layers)
    sd["token_embedding.weight"] = r.normal((a.vocab_size, dt), 0.02)
    sd["positional_embedding"] = r.normal((a.context_length, dt), 0.01)
    sd["ln_final.weight"] = 1.0 + r.normal((dt,), 0.1)
    sd["ln_final.bias"] = r.normal((dt,), 0.1)
    sd["text_projection"] = r.normal((dt, e), dt ** -0.5)
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07), dtype=torch.float32)
    return sd


def synthetic_prompt_params(arch: Arch, n_ctx: int, depth: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """The trainable tensors that do not depend on tokenisation (8 of the 10; `ctx` comes from
    the token embedding when CTX_INIT is given, trainers/mudpt.py:57-70).  Keys are relative
    to `visual.` / the prompt learner."""
    r = _Rng(seed + 7919)
    dv, dt, e = arch.vision_width, arch.transformer_width, arch.embed_dim
    out = {
        "visual.visual_ctx": r.normal((n_ctx, dv), 0.02),
        "visual.visual_ctx_deep_prompts": r.normal((depth - 1, n_ctx, dv), 0.02),
        "visual.visual_ctx_deep_projections.weight": r.uniform((e, dv), dv ** -0.5),
        "visual.visual_ctx_deep_projections.bias": r.uniform((e,), dv ** -0.5),
        "prompt.ctx_random": r.normal((n_ctx, dt), 0.02),
        "prompt.deep_prompts": r.normal((depth - 1, n_ctx, dt), 0.02),
        "prompt.embed_projection.weight": r.uniform((dv, dt), dt ** -0.5),
        "prompt.embed_projection.bias": r.uniform((dv,), dt ** -0.5),
        "prompt.deep_projections.weight": r.uniform((dv, dt), dt ** -0.5),
        "prompt.deep_projections.bias": r.uniform((dv,), dt ** -0.5),
    }
    return out


def synthetic_classnames(n_cls: int, kind: str = "short", seed: int = 0) -> List[str]:
    """`class i` names (EOT at 6/7/8 for 1/2/3 digits); `long` draws 1..8 extra words."""
    if kind == "short":
        return [f"class {i}" for i in range(n_cls)]
    g = np.random.default_rng(seed + 13)
    words = ["red", "small", "wild", "old", "northern", "spotted", "great", "common"]
    return [" ".join(list(g.choice(words, size=int(g.integers(1, 9)))) + [f"class {i}"]) for i in range(n_cls)]


def synthetic_images(batch: int, size: int = 224, seed: int = 1, kind: str = "noise") -> torch.Tensor:
    g = np.random.default_rng(seed + 104729)
    if kind == "noise":
        return torch.from_numpy(g.standard_normal((batch, 3, size, size), dtype=np.float32))
    if kind == "colour":  # per-image structure: raises logit margins
        c = torch.from_numpy(g.standard_normal((batch, 3, 1, 1), dtype=np.float32))
        return c.expand(batch, 3, size, size).contiguous()
    raise ValueError(kind)


def synthetic_labels(batch: int, n_cls: int, seed: int = 1) -> torch.Tensor:
    g = np.random.default_rng(seed + 1299709)
    return torch.from_numpy(g.integers(0, n_cls, size=(batch,), dtype=np.int64))


def assemble_state_dict(arch: Arch, tokenized_prompts: torch.Tensor, n_ctx: int, depth: int,
                        ctx_init_tokens=None, seed: int = 0, clip_sd=None) -> Dict[str, torch.Tensor]:
    """Flat fp32 state dict under the reference `CustomCLIP.state_dict()` names
    (`image_encoder.*`, `text_encoder.*`, `mudpt_prompt_learner.*`, `logit_scale`) built from
    the synthetic CLIP weights, the synthetic prompt parameters and a token-id matrix.

    `ctx_init_tokens`: token ids of CTX_INIT (row of clip.tokenize) -> ctx = embedding rows
    1..n_ctx (trainers/mudpt.py:57-63); None -> random ctx (:66-70)."""
    clip_sd = clip_sd if clip_sd is not None else synthetic_clip_state_dict(arch, seed)
    pp = synthetic_prompt_params(arch, n_ctx, depth, seed)
    emb = clip_sd["token_embedding.weight"]
    sd: Dict[str, torch.Tensor] = {}
    for k, v in clip_sd.items():
        if k.startswith("visual."):
            sd["image_encoder." + k[len("visual."):]] = v
        elif k.startswith("transformer.") or k in ("positional_embedding", "ln_final.weight",
                                                   "ln_final.bias", "text_projection"):
            sd["text_encoder." + k] = v
        elif k == "logit_scale":
            sd[k] = v
    for k, v in pp.items():
        if k.startswith("visual."):
            sd["image_encoder." + k[len("visual."):]] = v
        elif k != "prompt.ctx_random":
            sd["mudpt_prompt_learner." + k[len("prompt."):]] = v
    if ctx_init_tokens is not None:
        ids = torch.as_tensor(ctx_init_tokens).long().flatten()
        sd["mudpt_prompt_learner.ctx"] = emb[ids[1:1 + n_ctx]].clone()
    else:
        sd["mudpt_prompt_learner.ctx"] = pp["prompt.ctx_random"]
    tok = tokenized_prompts.long()
    e = emb[tok]  # [C, 77, dt]
    sd["mudpt_prompt_learner.token_prefix"] = e[:, :1, :].clone()
    sd["mudpt_prompt_learner.token_suffix"] = e[:, 1 + n_ctx:, :].clone()
    return sd


# ------------------------------------------------------------------------------------------------
# CoCoOp (BASELINE config 4, trainers/cocoop.py): plain CLIP towers + an instance-conditioned prompt learner
# ------------------------------------------------------------------------------------------------

def synthetic_cocoop_params(arch: Arch, n_ctx: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """ctx (random branch, trainers/cocoop.py:90-93) and the meta-net Linear(vis_dim, vis_dim // 16) ->
    ReLU -> Linear(vis_dim // 16, ctx_dim) (:99-103), torch default init ranges."""
    r = _Rng(seed + 15485863)
    dt, e = arch.transformer_width, arch.embed_dim
    hid = e // 16
    return {
        "ctx_random": r.normal((n_ctx, dt), 0.02),
        "meta_net.linear1.weight": r.uniform((hid, e), e ** -0.5), "meta_net.linear1.bias": r.uniform((hid,), e ** -0.5),
        "meta_net.linear2.weight": r.uniform((dt, hid), hid ** -0.5), "meta_net.linear2.bias": r.uniform((dt,), hid ** -0.5),
    }


def assemble_cocoop_state_dict(arch: Arch, tokenized_prompts: torch.Tensor, n_ctx: int, ctx_init_tokens=None,
                               seed: int = 0, clip_sd=None) -> Dict[str, torch.Tensor]:
    """Flat fp32 state dict under the reference CoCoOp `CustomCLIP.state_dict()` names
    (`image_encoder.*`, `text_encoder.*`, `prompt_learner.*`, `logit_scale`; trainers/cocoop.py:166-174)."""
    clip_sd = clip_sd if clip_sd is not None else synthetic_clip_state_dict(arch, seed)
    pp = synthetic_cocoop_params(arch, n_ctx, seed)
    emb = clip_sd["token_embedding.weight"]
    sd: Dict[str, torch.Tensor] = {}
    for k, v in clip_sd.items():
        if k.startswith("visual."):
            sd["image_encoder." + k[len("visual."):]] = v
        elif k.startswith("transformer.") or k in ("positional_embedding", "ln_final.weight", "ln_final.bias",
                                                   "text_projection"):
            sd["text_encoder." + k] = v
        elif k == "logit_scale":
            sd[k] = v
    for k, v in pp.items():
        if k != "ctx_random":
            sd["prompt_learner." + k] = v
    if ctx_init_tokens is not None:
        ids = torch.as_tensor(ctx_init_tokens).long().flatten()
        sd["prompt_learner.ctx"] = emb[ids[1:1 + n_ctx]].clone()
    else:
        sd["prompt_learner.ctx"] = pp["ctx_random"]
    e = emb[tokenized_prompts.long()]
    sd["prompt_learner.token_prefix"] = e[:, :1, :].clone()
    sd["prompt_learner.token_suffix"] = e[:, 1 + n_ctx:, :].clone()
    return sd


# ------------------------------------------------------------------------------------------------
# UMuDPT / UUMuDPT: the trainable tensors are many small modules (LightTransformer, LayerNorms,
# projections), so they are generated per NAME: both the reference model (oracle/make_golden.py) and the model
# under test load `param_by_name(name, shape)` into every trainable tensor.
# ------------------------------------------------------------------------------------------------

def param_by_name(name: str, shape, seed: int = 0) -> torch.Tensor:
    import zlib
    r = _Rng((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))
    shape = tuple(shape)
    leaf = name.split(".")[-1]
    is_ln = any(part.startswith("ln_") or "_ln_" in part for part in name.split(".")[:-1])
    if len(shape) == 1:
        if is_ln and leaf == "weight":
            return 1.0 + r.normal(shape, 0.1)
        return r.normal(shape, 0.05)
    if "ctx" in leaf or "prompts" in leaf:
        return r.normal(shape, 0.02)
    return r.normal(shape, float(shape[-1]) ** -0.5)


def make_cfg(n_ctx, depth, ctx_init, size, arch_name="ViT-B/16"):
    """The subset of the yacs tree the hot path reads (train.py:114-119); plain attribute dicts."""
    class N(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e
        __setattr__ = dict.__setitem__
    cfg = N()
    cfg.TRAINER = N(NAME="MuDPT", MUDPT=N(N_CTX=n_ctx, CTX_INIT=ctx_init, DEEP_PROMPT_DEPTH=depth, PREC="fp32"))
    cfg.INPUT = N(SIZE=(size, size))
    cfg.MODEL = N(BACKBONE=N(NAME=arch_name, PATH=""), INIT_WEIGHTS="")
    cfg.OPTIM = N(LR=0.0025, MAX_EPOCH=10)
    return cfg
