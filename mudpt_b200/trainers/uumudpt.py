"""UUMuDPT on the native towers -- drop-in for the reference's trainers/uumudpt.py.
UMuDPT plus vision-side prompts: `visual_ctx`, `visual_ctx_deep_prompts` and their LightTransformer, whose
output is added to the text deep prompts (clip/model.py:600-664, trainers/uumudpt.py:217-234)."""
from __future__ import annotations

import torch

from ..clip.model import _stack_with_ln_pre
from . import umudpt as _u
from .mudpt import TRAINER_REGISTRY


class UUMuDPTPromptLearner(_u.UMuDPTPromptLearner):
    CFG_NODE = "UUMUDPT"


class CustomCLIP(_u.CustomCLIP):
    LEARNER_ATTR = "uumudpt_prompt_learner"
    LEARNER_CLS = UUMuDPTPromptLearner

    def prompt_stacks(self):
        pl, ve = self.mudpt_prompt_learner, self.image_encoder
        vp = pl.visual_prompts()
        shared = vp[:1] + ve.visual_ctx.unsqueeze(0)                    # clip/model.py:636-637
        deeper = vp[1:] + ve.visual_ctx_deep_prompts                    # :641
        P_v = _stack_with_ln_pre(ve, shared, deeper)
        text_deep = pl.deep_prompts + ve.textual_prompts()             # trainers/uumudpt.py:224
        pos = self.text_encoder.positional_embedding[1:1 + pl.n_ctx]
        P_t = torch.cat([(pl.ctx + pos).unsqueeze(0), text_deep], dim=0).float()
        return P_v, P_t


@TRAINER_REGISTRY.register()
class UUMuDPT(_u.UMuDPT):
    MODEL_NAME = "UnifiedMultimodalDeepPromptTuning"
    CFG_NODE = "UUMUDPT"
    KEEP_VISUAL_CTX = True   # trainers/uumudpt.py:256-261
    CUSTOM_CLIP = CustomCLIP
