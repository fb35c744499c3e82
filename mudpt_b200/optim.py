"""Fused SGD for the handful of trainable prompt tensors: `torch.optim.SGD` semantics, one native
kernel launch per step (mudpt_sgd_step, include/mudpt_b200.h) instead of torch's per-operation foreach
launches.  This is the optimizer Dassl's `build_optimizer` creates for the MuDPT yamls (OPTIM.NAME = "sgd",
configs/trainers/MuDPT/*.yaml:15-22) and that `model_backward_and_update` steps (trainers/mudpt.py:251).

The class derives from torch.optim.SGD, so param_groups / state_dict / lr schedulers / checkpoints are
unchanged (state[p]["momentum_buffer"] keeps its name and meaning).  No CPU fallback: parameters that
are not contiguous fp32 CUDA tensors raise.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class FusedSGD(torch.optim.SGD):
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            if group.get("maximize", False):
                raise RuntimeError("FusedSGD: maximize is not supported")
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            momentum = float(group["momentum"])
            # torch initialises a momentum buffer on a parameter's first step (buf = clone(d_p)); the
            # trainable set steps together, so one first-step flag per group suffices
            new = [p for p in params if momentum != 0 and "momentum_buffer" not in self.state[p]]
            if new and len(new) != len(params):
                raise RuntimeError("FusedSGD: parameters of a group must start stepping together")
            bufs = []
            for p in params:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                        and p.grad.dtype == torch.float32):
                    raise RuntimeError("FusedSGD: contiguous fp32 CUDA parameters only (no CPU fallback)")
                if momentum != 0:
                    st = self.state[p]
                    if "momentum_buffer" not in st:
                        st["momentum_buffer"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                    bufs.append(st["momentum_buffer"])
            for lo in range(0, len(params), 32):
                chunk = params[lo:lo + 32]
                n = len(chunk)
                arr = C.c_void_p * n
                pp = arr(*[p.data_ptr() for p in chunk])
                gg = arr(*[p.grad.data_ptr() for p in chunk])
                bb = arr(*[b.data_ptr() for b in bufs[lo:lo + 32]]) if momentum != 0 else arr(*([None] * n))
                nn_ = (C.c_int64 * n)(*[p.numel() for p in chunk])
                _lib.check(lib.mudpt_sgd_step(pp, gg, bb, nn_, n, float(group["lr"]), momentum, float(group["dampening"]),
                                              float(group["weight_decay"]), 1 if group["nesterov"] else 0, 1 if new else 0,
                                              _lib.stream_ptr(chunk[0].device)))
        return loss
