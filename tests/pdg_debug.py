"""Test/debug access to the saved forward state (used by tests/ only)."""
from __future__ import annotations

import ctypes as C

import torch

from pdivgnn_b200 import _lib
from pdivgnn_b200.autograd import _params_struct, build_plan, param_list

WHAT = dict(x=0, e=1, y2=2, Pa=3, Pb=4, aggraw=5, hq=6, y3=7, y_nenc=8, y_eenc=9, hd=10, parts=11)


class ForwardState:
    def __init__(self, out, ws, plan, steps, flags):
        self.out, self.ws, self.plan, self.steps, self.flags = out, ws, plan, steps, flags

    def tensor(self, what: str, t: int = 0) -> torch.Tensor:
        L = _lib.lib()
        off, n = C.c_size_t(), C.c_size_t()
        _lib.check(L.pdg_ws_offset(self.plan.n_nodes, self.plan.n_edges, self.steps, self.flags, WHAT[what], t,
                                   C.byref(off), C.byref(n)), "pdg_ws_offset")
        if what == "parts":
            return self.ws[off.value:off.value + n.value * 8].view(torch.float64).view(-1, 2)
        return self.ws[off.value:off.value + n.value * 4].view(torch.float32).view(-1, 128)


def forward_with_state(model, batch, scale_output=False, scale_input=True) -> ForwardState:
    """pdg_forward with PDG_FLAG_SAVE, returning handles to every saved tensor."""
    L = _lib.lib()
    dev = batch.mean_stress.device
    n = batch.mean_stress.shape[0]
    plan = build_plan(batch.edge_index, n)
    flags = _lib.FLAG_SAVE | (_lib.FLAG_SCALE_INPUT if scale_input else 0) | (_lib.FLAG_SCALE_OUTPUT if scale_output else 0)
    steps = model.message_passing_steps
    params = [p.detach().contiguous() for p in param_list(model)]
    ps = _params_struct(params)
    norm = model._norm_struct()
    ws_bytes = L.pdg_forward_ws_bytes(n, plan.n_edges, steps, flags)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
    out = torch.empty((n, 3), dtype=torch.float32, device=dev)
    _lib.check(L.pdg_forward(C.byref(ps), C.byref(norm), _lib.ptr(batch.mean_stress.contiguous()),
                             _lib.ptr(batch.pos.contiguous()), _lib.ptr(batch.nodes_types.reshape(-1).contiguous()),
                             _lib.ptr(batch.edge_attr.contiguous()), _lib.ptr(plan.buf), n, plan.n_edges, steps, flags, 0,
                             _lib.ptr(ws), ws_bytes, _lib.ptr(out), _lib.stream_ptr(dev)), "pdg_forward")
    return ForwardState(out, ws, plan, steps, flags)
