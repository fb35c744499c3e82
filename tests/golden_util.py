"""Load a golden fixture (made from the reference by oracle/make_golden.py) and rebuild the
exact synthetic state dict it was produced with."""
import os

import numpy as np
import torch

from mudpt_b200 import synthetic as syn

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TINY = ["tiny_a", "tiny_b", "tiny_c", "tiny_d"]

_cache = {}


def load(name):
    if name in _cache:
        return _cache[name]
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    arch = syn.ARCHS[str(g["arch"])]
    n_ctx, depth, batch = int(g["n_ctx"]), int(g["depth"]), int(g["batch"])
    tok = torch.from_numpy(g["tokenized_prompts"])
    ctx_tokens = torch.from_numpy(g["ctx_init_tokens"]) if bool(g["has_ctx_init"]) else None
    sd = syn.assemble_state_dict(arch, tok, n_ctx, depth, ctx_tokens, seed=0)
    image = syn.synthetic_images(batch, arch.image_resolution, seed=1, kind=str(g["kind"]))
    labels = torch.from_numpy(g["labels"])
    case = dict(arch=arch, n_ctx=n_ctx, depth=depth, batch=batch, sd=sd, tokenized=tok, image=image,
                labels=labels, golden=g, classnames=[str(c) for c in g["classnames"]])
    _cache[name] = case
    return case
