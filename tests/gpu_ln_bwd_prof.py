"""LayerNorm backward on the bf16 gradient stream: us per launch of the kernel selected by MUDPT_LN_BWD_PIPE (0 register
kernel, 1 shared-memory pipeline, 2 pipeline with two rows per warp) at the tower shapes, against the HBM time of its
algorithmic bytes.  Run on the B200 box, one process per setting:
    MUDPT_LN_BWD_PIPE=1 python tests/gpu_ln_bwd_prof.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mudpt_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda")
st = _lib.stream_ptr(dev)
HBM = 6.55e12
out = {"MUDPT_LN_BWD_PIPE": os.environ.get("MUDPT_LN_BWD_PIPE", "default"), "MUDPT_LN_PIPE_STAGES": os.environ.get("MUDPT_LN_PIPE_STAGES"),
       "MUDPT_LN_PIPE_CTAS": os.environ.get("MUDPT_LN_PIPE_CTAS"), "shapes": {}}
# (name, sequences, L, row0, n, width, x as bf16 + statistics, residual: 1 bf16 in place / 2 fp32)
cases = [("text N=1 (1000 x 77, 512)", 1000, 77, 1, 2, 512, True, 1), ("text per rank (125 x 77, 512)", 125, 77, 1, 2, 512, True, 1),
         ("vision (32 x 199, 768) fp32 x", 32, 199, 197, 2, 768, False, 1), ("vision (32 x 199, 768) bf16 x", 32, 199, 197, 2, 768, True, 1),
         ("text N=1, fp32 residual in", 1000, 77, 1, 2, 512, True, 2), ("ViT-L vision (32 x 259, 1024) fp32 x", 32, 259, 257, 2, 1024, False, 1)]
for name, S, L, row0, n, d, xs, rk in cases:
    M = S * L
    nset = 3 if M * d * 8 > 40e6 else 1
    sets = []
    for _ in range(nset):
        x = torch.randn(M, d, device=dev)
        xb = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
        stats = torch.empty(M, d // 64, 2, device=dev)
        _lib.check(lib.mudpt_rowstats(x.data_ptr(), xb.data_ptr(), stats.data_ptr(), M, d, st))
        sets.append(dict(x=x, xb=xb, stats=stats, dy=torch.randn(M, d, device=dev).bfloat16(), g=torch.randn(d, device=dev),
                         r32=torch.randn(M, d, device=dev), rb=torch.randn(M, d, device=dev).bfloat16(), dx=torch.zeros(M, d, device=dev)))

    def run(b):
        _lib.check(lib.mudpt_layernorm_backward_stream(b["dy"].data_ptr(), b["xb"].data_ptr() if xs else b["x"].data_ptr(),
                                                       b["stats"].data_ptr() if xs else None, b["g"].data_ptr(),
                                                       b["rb"].data_ptr() if rk == 1 else b["r32"].data_ptr(), 1 if rk == 1 else 0,
                                                       b["dx"].data_ptr(), b["rb"].data_ptr(), M, d, L, row0, n, st))
    for b in sets:
        run(b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(24):
            run(sets[i % nset])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / 24)
    nbytes = M * d * (2 + (2 if xs else 4) + (2 if rk == 1 else 4) + 2)
    out["shapes"][name] = {"us": round(best, 2), "algorithmic_MB": round(nbytes / 1e6, 1), "hbm_us": round(nbytes / HBM * 1e6, 2),
                           "frac_of_hbm": round(nbytes / HBM * 1e6 / best, 3)}
print(json.dumps(out))
