"""Fused Adam for the 28 ``EncodeProcessDecode`` parameters.

Drop-in for ``torch.optim.Adam(model.parameters(), lr=...)`` in the reference train loop
(scripts/gnn_train.py:111,118,204-207): same update rule and ``state_dict`` layout
(``state[i] = {step, exp_avg, exp_avg_sq}``, one ``param_group``), but ONE kernel launch per
step (``pdg_adam_step``) instead of torch's ~12 foreach launches, and an optional
GradScaler-equivalent "skip the step on inf/nan gradients" decided on the device
(``pdg_grads_check_finite``), so ``scaler.step(optimizer)`` needs no host sync either.  With ``check_finite`` the
step count lives on the device too and advances only when the update is applied -- ``GradScaler.step`` does not
call ``optimizer.step()`` on overflow (gnn_train.py:205-207), so bias corrections and the saved ``state['step']``
stay identical to the reference's after a skipped step.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 check_finite: bool = False):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam takes the model's parameters as one group")
        ps = self.param_groups[0]["params"]
        if len(ps) != _lib.PDG_NUM_PARAMS or sum(p.numel() for p in ps) != _lib.PDG_PARAM_ELEMS:
            raise ValueError("FusedAdam expects the 28 EncodeProcessDecode parameters in state_dict order")
        for p in ps:
            _lib.require_cuda(p, "parameter", torch.float32)
            if not p.is_contiguous():
                raise ValueError("parameters must be contiguous")
        dev = ps[0].device
        self.check_finite = check_finite
        self._exp_avg = torch.zeros(_lib.PDG_PARAM_ELEMS, dtype=torch.float32, device=dev)
        self._exp_avg_sq = torch.zeros_like(self._exp_avg)
        self._found_inf = torch.zeros(1, dtype=torch.int32, device=dev)
        self._steps = 0      # host count of step() calls (== applied updates when check_finite is off)
        self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev)  # applied updates (check_finite: the truth)
        self._bind_state()

    # state exposed exactly like torch.optim.Adam (views into the two flat moment buffers)
    def _bind_state(self):
        off = 0
        for p in self.param_groups[0]["params"]:
            n = p.numel()
            self.state[p] = {"step": torch.tensor(float(self.applied_steps())),
                             "exp_avg": self._exp_avg[off:off + n].view(p.shape),
                             "exp_avg_sq": self._exp_avg_sq[off:off + n].view(p.shape)}
            off += n

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)  # replaces the state tensors by copies: fold them back into the flat buffers
        ps = self.param_groups[0]["params"]
        off, steps = 0, 0
        for p in ps:
            st = self.state.get(p, {})
            n = p.numel()
            if "exp_avg" in st:
                self._exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self._exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps = int(float(st["step"]))
            off += n
        self._steps = steps
        self._step_dev.fill_(steps)
        self._bind_state()

    def applied_steps(self) -> int:
        """Number of updates actually applied.  Without ``check_finite`` this is the host-side call count; with it
        the count lives on the device (skipped steps do not advance it) and reading it is a device->host copy."""
        return int(self._step_dev.item()) if self.check_finite else self._steps

    def state_dict(self):
        if self.check_finite:  # refresh the exposed per-parameter `step` tensors from the device count (checkpoint time)
            n = float(self.applied_steps())
            for p in self.param_groups[0]["params"]:
                self.state[p]["step"] = torch.tensor(n)
        return super().state_dict()

    @property
    def found_inf(self) -> torch.Tensor:
        """Device int: != 0 when the last ``step(check_finite=True)`` was skipped."""
        return self._found_inf

    @torch.no_grad()
    def step(self, closure=None, inv_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        g = self.param_groups[0]
        ps = g["params"]
        arr_p, arr_g = (C.c_void_p * _lib.PDG_NUM_PARAMS)(), (C.c_void_p * _lib.PDG_NUM_PARAMS)()
        for i, p in enumerate(ps):
            if p.grad is None:
                raise RuntimeError("FusedAdam.step: every parameter needs a gradient (the fused backward always "
                                   "produces all 28)")
            gr = p.grad
            if not gr.is_cuda or gr.dtype != torch.float32 or not gr.is_contiguous():
                raise RuntimeError("FusedAdam.step: gradients must be contiguous CUDA float32 tensors")
            arr_p[i], arr_g[i] = p.data_ptr(), gr.data_ptr()
        dev = ps[0].device
        self._steps += 1
        cfg = _lib.PdgAdam(lr=float(g["lr"]), beta1=float(g["betas"][0]), beta2=float(g["betas"][1]), eps=float(g["eps"]),
                           weight_decay=float(g["weight_decay"]), inv_scale=float(inv_scale), step=self._steps)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            if self.check_finite:
                self._found_inf.zero_()
                _lib.check(L.pdg_grads_check_finite(C.byref(arr_g), _lib.ptr(self._found_inf), st),
                           "pdg_grads_check_finite")
                # device-side step count: a skipped step leaves it (and the next bias corrections) untouched
                _lib.check(L.pdg_adam_step_counted(C.byref(arr_p), C.byref(arr_g), _lib.ptr(self._exp_avg),
                                                   _lib.ptr(self._exp_avg_sq), C.byref(cfg), _lib.ptr(self._found_inf),
                                                   _lib.ptr(self._step_dev), st), "pdg_adam_step_counted")
            else:
                _lib.check(L.pdg_adam_step(C.byref(arr_p), C.byref(arr_g), _lib.ptr(self._exp_avg),
                                           _lib.ptr(self._exp_avg_sq), C.byref(cfg), None, st), "pdg_adam_step")
        if not self.check_finite:
            for p in ps:
                self.state[p]["step"] += 1
        return loss
