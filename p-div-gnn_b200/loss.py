"""Fused training loss (per-graph NMSE + divergence regulariser) on the device.

Replaces the per-graph Python loop of the reference ``train()`` (gnn_train.py:162-202)
-- ``data_utils.slice_batch_gt_and_predictions`` (data_utils.py:25-33),
``normalized_mse_loss_single`` (gnn_train.py:41-57), ``compute_divergence`` (:60-92),
``data_utils.standardize`` (data_utils.py:46-51) -- by two kernel launches that work on
the batch's ``ptr`` segments.  Returns the two 0-dim device tensors the reference logs
separately (gnn_train.py:194-202): ``nmse`` and ``divergence`` (already multiplied by
``divergence_penalty`` and divided by the number of graphs); the training loss is their sum.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch

from . import _lib

_OP_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()


class OpDivPlan:
    __slots__ = ("buf", "values", "nnz", "n_nodes", "keep")

    def __init__(self, buf, values, nnz, n_nodes, keep):
        self.buf, self.values, self.nnz, self.n_nodes, self.keep = buf, values, nnz, n_nodes, keep


def build_opdiv_plan(op_div_matrix: torch.Tensor, graph_ptr: torch.Tensor) -> OpDivPlan:
    """CSR + CSC of the row-stacked sparse operator of a PyG batch."""
    L = _lib.lib()
    if not op_div_matrix.is_sparse:
        raise TypeError("op_div_matrix must be a torch sparse COO tensor")
    if not op_div_matrix.is_cuda:
        raise RuntimeError("op_div_matrix must live on the GPU; there is no CPU path")
    key = (id(op_div_matrix), graph_ptr.data_ptr(), graph_ptr._version)
    hit = _OP_CACHE.get(key)
    if hit is not None:
        return hit
    m = op_div_matrix if op_div_matrix.is_coalesced() else op_div_matrix.coalesce()
    idx = m.indices().contiguous()
    val = _lib.require_cuda(m.values(), "op_div values", torch.float32)
    nnz, n = idx.shape[1], m.shape[0]
    dev = val.device
    gptr = _lib.require_cuda(graph_ptr, "ptr", torch.int64)
    with torch.cuda.device(dev):
        buf = torch.empty(L.pdg_opdiv_plan_bytes(n, nnz), dtype=torch.uint8, device=dev)
        tb = L.pdg_opdiv_tmp_bytes(n, nnz)
        tmp = torch.empty(tb, dtype=torch.uint8, device=dev)
        _lib.check(L.pdg_opdiv_plan_build(_lib.ptr(idx[0]), _lib.ptr(idx[1]), nnz, _lib.ptr(gptr), gptr.numel() - 1, n,
                                          _lib.ptr(buf), _lib.ptr(tmp), tb, _lib.stream_ptr(dev)),
                   "pdg_opdiv_plan_build")
    plan = OpDivPlan(buf, val, nnz, n, (op_div_matrix, graph_ptr))
    _OP_CACHE[key] = plan
    while len(_OP_CACHE) > 8:
        _OP_CACHE.popitem(last=False)
    return plan


class _LossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, local_stress, gptr, labels, plan, norm, use_div, penalty):
        L = _lib.lib()
        dev = pred.device
        n, b = pred.shape[0], gptr.numel() - 1
        with torch.cuda.device(dev):
            ws = torch.empty(L.pdg_loss_ws_bytes(n, b), dtype=torch.uint8, device=dev)
            out2 = torch.empty(2, dtype=torch.float32, device=dev)
            _lib.check(L.pdg_loss(_lib.ptr(pred), _lib.ptr(local_stress), C.byref(norm), _lib.ptr(gptr), b, n,
                                  _lib.ptr(labels), _lib.ptr(plan.buf) if plan else None,
                                  _lib.ptr(plan.values) if plan else None, plan.nnz if plan else 0, int(use_div),
                                  float(penalty), _lib.ptr(ws), _lib.ptr(out2), _lib.stream_ptr(dev)), "pdg_loss")
        ctx.pdg = (pred, local_stress, gptr, plan, norm, use_div, penalty, ws)
        return out2[0], out2[1]

    @staticmethod
    def backward(ctx, g_nmse, g_div):
        L = _lib.lib()
        pred, local_stress, gptr, plan, norm, use_div, penalty, ws = ctx.pdg
        dev = pred.device
        n, b = pred.shape[0], gptr.numel() - 1
        with torch.cuda.device(dev):
            up = torch.stack([g_nmse.reshape(()), g_div.reshape(())]).to(torch.float32).contiguous()
            grad = torch.empty_like(pred)
            _lib.check(L.pdg_loss_backward(_lib.ptr(pred), _lib.ptr(local_stress), C.byref(norm), _lib.ptr(gptr), b, n,
                                           _lib.ptr(plan.buf) if plan else None,
                                           _lib.ptr(plan.values) if plan else None, plan.nnz if plan else 0,
                                           int(use_div), float(penalty), _lib.ptr(ws), _lib.ptr(up), _lib.ptr(grad),
                                           _lib.stream_ptr(dev)), "pdg_loss_backward")
        return grad, None, None, None, None, None, None, None


def nmse_div_loss(predicted_local_stress: torch.Tensor, mesh_graph_batch, model, optimize_divergence: bool = True,
                  divergence_penalty: float = 1.0):
    """(nmse, divergence) of one batch; ``loss = nmse + divergence`` (gnn_train.py:193-202).

    ``predicted_local_stress`` is the standardised prediction (``scale_output=False``),
    ``mesh_graph_batch.local_stress`` the RAW ground truth (standardised inside with the
    model's ``mean/std_local_stress``, gnn_train.py:162-167).
    """
    pred = _lib.require_cuda(predicted_local_stress, "predicted_local_stress", torch.float32)
    ls = _lib.require_cuda(mesh_graph_batch.local_stress, "local_stress", torch.float32)
    gptr = getattr(mesh_graph_batch, "ptr", None)
    if gptr is None:  # un-batched Data (benchmark_gnn_fem.py:97): one graph
        gptr = torch.tensor([0, pred.shape[0]], dtype=torch.int64, device=pred.device)
    gptr = _lib.require_cuda(gptr, "ptr", torch.int64)
    labels, plan = None, None
    if optimize_divergence:
        labels = _lib.require_cuda(mesh_graph_batch.surfaces_nodes_for_div, "surfaces_nodes_for_div", torch.int64).reshape(-1)
        plan = build_opdiv_plan(mesh_graph_batch.op_div_matrix, gptr)
    return _LossFunction.apply(pred, ls, gptr, labels, plan, model._norm_struct(), bool(optimize_divergence),
                               float(divergence_penalty))
