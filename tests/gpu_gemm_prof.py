"""Profiling / timing driver for single GEMM shapes with their fused epilogues (run on the B200 box).

    python tests/gpu_gemm_prof.py [tag ...] [--iters N] [--m ROWS]

tags: qkv out c_fc c_proj d_c_proj d_c_fc d_out d_qkv  (text tower, d=512)  or v_<tag> (vision tower, d=768).
Prints one line per tag with CUDA-event time; under `ncu -k regex:gemm` the same launches are captured.
Test infrastructure only.
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def shapes(d):
    # tag -> (N, K, mode)
    return {"qkv": (3 * d, d, 0), "out": (d, d, 2), "c_fc": (4 * d, d, 3), "c_proj": (d, 4 * d, 2),
            "d_c_proj": (4 * d, d, 4), "d_c_fc": (d, 4 * d, 0), "d_out": (d, d, 0), "d_qkv": (d, 3 * d, 0)}


def main():
    import torch
    from mudpt_b200 import _lib
    args = sys.argv[1:]
    iters, rows = 10, None
    tags = []
    i = 0
    while i < len(args):
        if args[i] == "--iters":
            iters = int(args[i + 1]); i += 2
        elif args[i] == "--m":
            rows = int(args[i + 1]); i += 2
        else:
            tags.append(args[i]); i += 1
    tags = tags or ["qkv", "out", "c_fc", "c_proj", "d_c_proj", "d_c_fc", "d_out", "d_qkv"]
    lib = _lib.load()
    dev = torch.device("cuda")
    st = _lib.stream_ptr(dev)
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for tag in tags:
        vis = tag.startswith("v_")
        d = 768 if vis else 512
        M = rows or (6368 if vis else 77000)
        N, K, mode = shapes(d)[tag[2:] if vis else tag]
        A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
        bias = torch.randn(N, device=dev)
        f32 = mode in (1, 2)
        out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
        out1 = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if mode == 3 else None
        resid = torch.randn(M, N, device=dev) if mode == 2 else None
        aux = torch.randn(M, N, device=dev).bfloat16() if mode == 4 else None

        def call():
            _lib.check(lib.mudpt_gemm_bf16(A.data_ptr(), B.data_ptr(), M, N, K, mode, out.data_ptr(),
                                           out1.data_ptr() if out1 is not None else None, bias.data_ptr(),
                                           resid.data_ptr() if resid is not None else None,
                                           aux.data_ptr() if aux is not None else None, N, 1, 1, st))
        for _ in range(2):
            call()
        ts = []
        for _ in range(iters):
            flush.zero_()  # L2 flush between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        nbytes = 2 * (M * K + N * K) + M * N * ((4 if f32 else 2) + (2 if mode == 3 else 0) + (4 if mode == 2 else 0) + (2 if mode == 4 else 0))
        print(f"{tag}: M={M} N={N} K={K} mode={mode}  {ms * 1e3:.1f} us  {2 * M * N * K / ms / 1e9:.0f} TFLOP/s  "
              f"{nbytes / ms / 1e6:.0f} GB/s  (hbm-bound {nbytes / 6538.3e3:.1f} us, mma-bound {2 * M * N * K / 1618.5e6:.1f} us)", flush=True)


if __name__ == "__main__":
    main()
