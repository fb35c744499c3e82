"""Timing / profiling driver for the attention kernels at the BASELINE config-2 shapes (B200 box).

    python tests/gpu_attn_prof.py [vision|text77|text9|text77_8 ...] [--prof]

--prof: 3 forward + 3 backward launches only (ncu: -k regex:attn -s 2 -c 3 captures fwd, dQ, dK/dV).
Test infrastructure only.
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {"vision": (32, 199, 12, 0), "text77": (1000, 77, 8, 1), "text9": (1000, 9, 8, 1), "text77_8": (125, 77, 8, 1),
          "vitl": (32, 259, 16, 0)}


def main():
    import torch
    from mudpt_b200 import _lib
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    prof = "--prof" in sys.argv
    lib = _lib.load()
    dev = torch.device("cuda")
    st = _lib.stream_ptr(dev)
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for tag in (args or ["vision", "text77", "text77_8", "text9"]):
        S, L, H, causal = SHAPES[tag]
        d = H * 64
        qkv = torch.randn(S * L, 3 * d, device=dev).bfloat16()
        o = torch.zeros(S * L, d, device=dev, dtype=torch.bfloat16); lse = torch.zeros(S, H, L, device=dev)
        do = torch.randn(S * L, d, device=dev).bfloat16(); dqkv = torch.zeros(S * L, 3 * d, device=dev, dtype=torch.bfloat16)
        dsum = torch.zeros(S, H, L, device=dev)
        fwd = lambda: _lib.check(lib.mudpt_attention_forward(qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), S, L, H, causal, st))
        bwd = lambda: _lib.check(lib.mudpt_attention_backward(qkv.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                                             dsum.data_ptr(), dqkv.data_ptr(), S, L, H, causal, st))
        for name, fn, mult in (("fwd", fwd, 4), ("bwd", bwd, 8)):
            if prof:
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                continue
            for _ in range(3):
                fn()
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            us = ts[len(ts) // 2] * 1e3
            nbytes = S * L * d * 2 * mult
            flops = 4.0 * S * H * L * L * 64 * (1.0 if name == "fwd" else 2.5)
            print(f"{tag} {name}: S={S} L={L} H={H}  {us:.1f} us  {flops / us / 1e6:.0f} TFLOP/s  {nbytes / us / 1e3:.0f} GB/s "
                  f"(hbm-bound {nbytes / 6538.3e3:.1f} us)", flush=True)


if __name__ == "__main__":
    main()
