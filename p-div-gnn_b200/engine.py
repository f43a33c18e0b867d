"""Training / evaluation / inference loops around the hot path.

Device-side counterparts of the reference drivers' inner loops:

  :func:`train_epoch`       scripts/gnn_train.py:152-207  (forward, fused per-graph loss, backward, optimizer step)
  :func:`evaluate`          scripts/gnn_train.py:208-253  (test NMSE (+ monitored divergence) per batch, no grad)
  :func:`predict_and_save`  scripts/gnn_inference.py:45-81 (un-standardised prediction, one ``.npz`` per mesh)

The per-graph Python loops of the reference (``slice_batch_gt_and_predictions``, ``slice_batch_predictions``,
``criterion`` per graph) are replaced by the segmented loss kernels and by ONE device->host copy per batch that
is split by the batch's ``ptr`` on the host.  Epoch sums stay on the device; ``.item()`` is called once per epoch.
"""
from __future__ import annotations

import os
import shutil
from typing import Optional

import numpy as np
import torch

from .loss import nmse_div_loss


def train_epoch(model, loader, optimizer, optimize_divergence: bool = True, divergence_penalty: float = 1.0) -> dict:
    """One pass over ``loader`` (gnn_train.py:152-207).  Returns the epoch means the reference logs
    (``Loss/Loss Train``, ``Loss/MSE Train``, ``Loss/Divergence Train``)."""
    model.train()
    dev = next(model.parameters()).device
    tot = torch.zeros(3, device=dev)
    nb = 0
    for batch in loader:
        pred = model(batch, scale_output=False, scale_input=True).local_stress
        nmse, div = nmse_div_loss(pred, batch, model, optimize_divergence, divergence_penalty)
        loss = nmse + div
        optimizer.zero_grad(set_to_none=True)
        loss.backward()
        optimizer.step()
        tot += torch.stack([loss.detach(), nmse.detach(), div.detach()])
        nb += 1
    t = (tot / max(nb, 1)).tolist()  # the epoch's only device->host read
    return {"total": t[0], "nmse": t[1], "divergence": t[2], "batches": nb}


@torch.no_grad()
def evaluate(model, loader, monitor_divergence: bool = False) -> dict:
    """Test pass (gnn_train.py:208-253): mean over batches of [sum_i NMSE_i / B (+ sum_i div_i / B)].

    As in the reference the monitored divergence is NOT multiplied by the training penalty
    (``compute_divergence(..., reduce_strategy="square")`` is added as is)."""
    model.eval()
    dev = next(model.parameters()).device
    tot = torch.zeros(3, device=dev)
    nb = 0
    for batch in loader:
        pred = model(batch, scale_output=False, scale_input=True).local_stress
        nmse, div = nmse_div_loss(pred, batch, model, monitor_divergence, 1.0)
        tot += torch.stack([nmse + div, nmse, div])
        nb += 1
    t = (tot / max(nb, 1)).tolist()
    return {"total": t[0], "nmse": t[1], "divergence": t[2], "batches": nb}


@torch.no_grad()
def predict(model, loader):
    """Yield ``(sample_id, stress_field [N_i,3] float32 numpy)`` in dataset order (``scale_output=True``)."""
    model.eval()
    for batch in loader:
        out = model(batch, scale_output=True, scale_input=True).local_stress
        host = out.cpu().numpy()  # one copy per batch; split by ptr on the host (slice_batch_predictions, data_utils.py:36-43)
        ptr = batch.ptr.cpu().numpy()
        ids = getattr(batch, "sample_ids", None) or list(range(len(ptr) - 1))
        for k, sid in enumerate(ids):
            yield sid, host[ptr[k]:ptr[k + 1]]


def predict_and_save(model, loader, results_folder: str, data_filenames: Optional[list] = None) -> list:
    """gnn_inference.py:45-81: ``fields/hole_plate_mesh_<id>.npz`` = a copy of the sample's original data file with
    ``stress_field`` replaced by the prediction.  The loader must not shuffle.  Returns the written paths."""
    fields = os.path.join(results_folder, "fields")
    os.makedirs(fields, exist_ok=True)
    if data_filenames is None:
        data_filenames = list(loader.ds.dataframe["data_filename"])
    written = []
    mesh_id = 0
    for sid, field in predict(model, loader):
        if sid != mesh_id:
            raise RuntimeError("predict_and_save needs an un-shuffled, un-sharded loader (THE DATASET MUST NOT BE SHUFFLED)")
        target = os.path.join(fields, f"hole_plate_mesh_{mesh_id}.npz")
        src = data_filenames[mesh_id]
        shutil.copyfile(src, target)
        org = dict(np.load(src))
        org["stress_field"] = field
        np.savez(target, **org)
        written.append(target)
        mesh_id += 1
    return written
