"""Device-side graph batcher front-end.

Turns a list of meshes (coordinates + triangles + per-node fields, e.g. the samples of
:mod:`synth` or the arrays of the reference's ``.vtk``/``.npz`` pairs) into ONE batched
graph on the GPU with the attribute names the reference's ``Batch`` carries
(datasets.py:240-281 + PyG collation): ``pos, edge_index, edge_attr, mean_stress,
local_stress, nodes_types, surfaces_nodes_for_div, op_div_matrix, batch, ptr``.
Edge construction (FaceToEdge + undirected + lengths + periodic edges + coalesce + node
offsets) runs in ``pdg_batch_count/fill``; the per-node fields are plain concatenations.
"""
from __future__ import annotations

import collections
import concurrent.futures
import ctypes as C

import numpy as np
import torch

from . import _lib


class MeshBatch:
    """Duck-typed ``torch_geometric.data.Batch`` (attribute access, ``.to``, ``len``)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(vars(self).items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self

    def __len__(self):
        return self.batch_size

    @property
    def num_graphs(self):
        return self.batch_size


def _nodes_per_face(faces: torch.Tensor) -> int:
    """3 (triangles) or 4 (quads): the reference dispatches on the first cell's type and assumes one type per mesh
    (convert_utils.py:24,52-58); a batch is one dataset, so one type per batch."""
    if faces.dim() != 2 or faces.shape[0] not in (3, 4):
        raise NotImplementedError(f"faces must be [3,F] (triangles) or [4,F] (quads); got {tuple(faces.shape)}")
    return int(faces.shape[0])


def build_edges(pos64: torch.Tensor, faces: torch.Tensor, node_ptr: torch.Tensor, face_ptr: torch.Tensor,
                periodic: bool = True):
    """(edge_index [2,E] int64, edge_attr [E] fp32) of B concatenated meshes, on the GPU.

    pos64 [N,2] float64; faces [3,F] (triangles, FaceToEdge) or [4,F] (quads, convert_utils._quad_face_to_edge:62-81)
    int64 graph-LOCAL node ids, one cell type per batch; node_ptr/face_ptr [B+1] int64.
    """
    L = _lib.lib()
    pos64 = _lib.require_cuda(pos64, "pos", torch.float64)
    faces = _lib.require_cuda(faces, "faces", torch.int64)
    node_ptr = _lib.require_cuda(node_ptr, "node_ptr", torch.int64)
    face_ptr = _lib.require_cuda(face_ptr, "face_ptr", torch.int64)
    npf = _nodes_per_face(faces)
    n, f, b = pos64.shape[0], faces.shape[1], node_ptr.numel() - 1
    dev = pos64.device
    with torch.cuda.device(dev):
        tb = L.pdg_batch_tmp_bytes(n, f, npf, b)
        tmp = torch.empty(tb, dtype=torch.uint8, device=dev)
        ne = C.c_int64(0)
        _lib.check(L.pdg_batch_count(_lib.ptr(pos64), _lib.ptr(faces), _lib.ptr(node_ptr), _lib.ptr(face_ptr), b, n, f,
                                     npf, int(periodic), _lib.ptr(tmp), tb, C.byref(ne), _lib.stream_ptr(dev)),
                   "pdg_batch_count")
        e = ne.value
        edge_index = torch.empty((2, e), dtype=torch.int64, device=dev)
        edge_attr = torch.empty(e, dtype=torch.float32, device=dev)
        _lib.check(L.pdg_batch_fill(_lib.ptr(pos64), n, f, npf, b, e, _lib.ptr(tmp), _lib.ptr(edge_index),
                                    _lib.ptr(edge_attr), _lib.stream_ptr(dev)), "pdg_batch_fill")
    return edge_index, edge_attr


def node_labels(pos64: torch.Tensor, faces: torch.Tensor, node_ptr: torch.Tensor, face_ptr: torch.Tensor,
                check_regions: bool = False):
    """``datasets.compute_node_labels`` (datasets.py:133-179) for B concatenated meshes, on the GPU.

    Returns ``(labels [N] int64 in {-1, 0, 1}, n_regions [B] int32)``.  ``check_regions`` repeats the reference's
    ``assert n_regions == 2`` (one device->host read).
    """
    L = _lib.lib()
    pos64 = _lib.require_cuda(pos64, "pos", torch.float64)
    faces = _lib.require_cuda(faces, "faces", torch.int64)
    node_ptr = _lib.require_cuda(node_ptr, "node_ptr", torch.int64)
    face_ptr = _lib.require_cuda(face_ptr, "face_ptr", torch.int64)
    if pos64.dim() != 2 or pos64.shape[1] != 2:
        raise ValueError("pos must be [N, 2] float64")
    npf = _nodes_per_face(faces)
    n, f, b = pos64.shape[0], faces.shape[1], node_ptr.numel() - 1
    dev = pos64.device
    with torch.cuda.device(dev):
        tb = L.pdg_labels_tmp_bytes(n, f, npf, b)
        tmp = torch.empty(tb, dtype=torch.uint8, device=dev)
        labels = torch.empty(n, dtype=torch.int64, device=dev)
        regions = torch.empty(b, dtype=torch.int32, device=dev)
        _lib.check(L.pdg_node_labels(_lib.ptr(pos64), _lib.ptr(faces), _lib.ptr(node_ptr), _lib.ptr(face_ptr), b, n, f,
                                     npf, _lib.ptr(tmp), tb, _lib.ptr(labels), _lib.ptr(regions), _lib.stream_ptr(dev)),
                   "pdg_node_labels")
    if check_regions:
        bad = (regions != 2).nonzero().flatten().tolist()
        if bad:
            raise AssertionError(f"Expected 2 regions, found {regions[bad[0]].item()} for mesh {bad[0]} of the batch")
    return labels, regions


def is_periodic(pos64: torch.Tensor, node_ptr: torch.Tensor = None, tol: float = 1e-8) -> torch.Tensor:
    """``microgen.mesh.is_periodic`` (asserted at generate_dataset.py:191, benchmark_gnn_fem.py:195) for B concatenated
    2-D meshes, on the GPU: bool tensor [B] (device; no host sync).  pos64 [N,2] float64 (a third column, the z = 0 of
    ``shape.points``, is dropped like the reference's ``[:, :-1]``); node_ptr [B+1] int64, default one mesh."""
    L = _lib.lib()
    if pos64.dim() == 2 and pos64.shape[1] == 3:
        pos64 = pos64[:, :2].contiguous()
    pos64 = _lib.require_cuda(pos64, "pos", torch.float64)
    if pos64.dim() != 2 or pos64.shape[1] != 2:
        raise ValueError("pos must be [N, 2] float64")
    dev = pos64.device
    n = pos64.shape[0]
    if node_ptr is None:
        node_ptr = torch.tensor([0, n], dtype=torch.int64, device=dev)
    node_ptr = _lib.require_cuda(node_ptr, "node_ptr", torch.int64)
    b = node_ptr.numel() - 1
    with torch.cuda.device(dev):
        tb = L.pdg_periodic_tmp_bytes(n, b)
        tmp = torch.empty(tb, dtype=torch.uint8, device=dev)
        flags = torch.empty(b, dtype=torch.int32, device=dev)
        _lib.check(L.pdg_is_periodic(_lib.ptr(pos64), _lib.ptr(node_ptr), b, n, float(tol), _lib.ptr(tmp), tb,
                                     _lib.ptr(flags), _lib.stream_ptr(dev)), "pdg_is_periodic")
    return flags != 0


def convert_mesh_to_graph(points, faces, mean_stress, device="cuda", check_periodic: bool = True) -> MeshBatch:
    """``benchmark_gnn_fem.convert_mesh_to_graph`` (benchmark_gnn_fem.py:388-415, the "with preprocessing" series of the
    reference's benchmark) with every step on the GPU: mesh -> sorted symmetric edges, edge lengths, periodic edges
    (``pdg_batch_count/fill``), node labels as ``surfaces_nodes_for_div`` / ``nodes_types`` (``pdg_node_labels``), 2-D
    fp32 positions and the broadcast mean stress.  ``points`` [N,2|3] float64 and ``faces`` [3,F] / [4,F] int64 may be
    host (numpy / CPU tensors) or CUDA tensors; ``check_periodic`` repeats the reference's
    ``assert is_periodic(shape.points[:, :-1])`` (benchmark_gnn_fem.py:195; one device->host read).
    The result is an un-batched graph (no ``batch`` / ``ptr``), what the benchmark feeds to ``model.forward``."""
    dev = torch.device(device)
    pos64 = torch.as_tensor(points, dtype=torch.float64)
    if pos64.dim() != 2 or pos64.shape[1] not in (2, 3):
        raise ValueError("points must be [N, 2] or [N, 3]")
    pos64 = pos64[:, :2].contiguous().to(dev, non_blocking=True)
    face = torch.as_tensor(faces, dtype=torch.int64).contiguous().to(dev, non_blocking=True)
    n, f = pos64.shape[0], face.shape[1]
    node_ptr = torch.tensor([0, n], dtype=torch.int64).to(dev, non_blocking=True)
    face_ptr = torch.tensor([0, f], dtype=torch.int64).to(dev, non_blocking=True)
    periodic_ok = is_periodic(pos64, node_ptr) if check_periodic else None
    edge_index, edge_attr = build_edges(pos64, face, node_ptr, face_ptr, periodic=True)
    labels, _ = node_labels(pos64, face, node_ptr, face_ptr)
    if check_periodic and not bool(periodic_ok.item()):
        raise AssertionError("Mesh is not periodic")
    ms = torch.as_tensor(tuple(float(v) for v in mean_stress), dtype=torch.float32).to(dev, non_blocking=True)
    lab = labels.unsqueeze(1)
    return MeshBatch(pos=pos64.float(), face=face, edge_index=edge_index, edge_attr=edge_attr,
                     mean_stress=ms.expand(n, 3).contiguous(), surfaces_nodes_for_div=lab, nodes_types=lab.clone(),
                     is_periodic=True, num_nodes=n, batch_size=1)


def host_arrays(samples, with_op_div: bool = True, pin: bool = True):
    """Concatenate a list of mesh samples (dicts, see synth.make_rve_mesh) into flat host arrays
    (pinned when CUDA is available and ``pin``) -- the layout a real data loader would hand to the GPU.
    ``with_op_div=False`` leaves the divergence operator out (inference / divergence-free training)."""
    ns = [s["pos"].shape[0] for s in samples]
    fs = [s["faces"].shape[1] for s in samples]
    nptr = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    fptr = np.concatenate([[0], np.cumsum(fs)]).astype(np.int64)
    rows, cols, vals = [], [], []
    for s, off in zip(samples if with_op_div else [], nptr[:-1]):
        r, c, v = np.asarray(s["op_div_row"]), np.asarray(s["op_div_col"]), np.asarray(s["op_div_data"])
        order = np.lexsort((c, r))  # coalesced order (row, col); entries are unique per sample
        rows.append(r[order] + off)
        cols.append(c[order])
        vals.append(v[order].astype(np.float32))
    h = dict(
        pos64=np.concatenate([np.asarray(s["pos"])[:, :2] for s in samples]).astype(np.float64),
        faces=np.concatenate([np.asarray(s["faces"]) for s in samples], axis=1).astype(np.int64),
        node_ptr=nptr, face_ptr=fptr,
        mean_stress=np.concatenate([np.ones((n, 3), np.float32) * np.asarray(s["mean_stress"], np.float32)[None, :]
                                    for s, n in zip(samples, ns)]),
        local_stress=np.concatenate([np.asarray(s["stress_field"]) for s in samples]).astype(np.float32),
        labels=np.concatenate([np.asarray(s["labels"]) for s in samples]).astype(np.int64),
    )
    if with_op_div:
        h.update(op_row=np.concatenate(rows).astype(np.int64), op_col=np.concatenate(cols).astype(np.int64),
                 op_val=np.concatenate(vals), op_width=np.array(max(2 * n for n in ns), dtype=np.int64))
    out = {}
    pin = pin and torch.cuda.is_available()
    for k, v in h.items():
        t = torch.from_numpy(np.ascontiguousarray(v))
        out[k] = t.pin_memory() if pin and t.dim() > 0 else t
    return out


def host_bytes(h, with_op_div: bool = True) -> int:
    """Bytes batch_from_host copies host -> device."""
    return int(sum(t.numel() * t.element_size() for k, t in h.items()
                   if k != "op_width" and (with_op_div or not k.startswith("op_"))))


def batch_from_host(h, device="cuda", periodic: bool = True, with_op_div: bool = True) -> MeshBatch:
    """H2D copies (non_blocking from pinned memory) + device edge construction -> MeshBatch."""
    if with_op_div and "op_row" not in h:
        raise ValueError("batch_from_host(with_op_div=True) needs host arrays built with the divergence operator")
    d = {k: v.to(device, non_blocking=True) for k, v in h.items()
         if k != "op_width" and (with_op_div or not k.startswith("op_"))}
    edge_index, edge_attr = build_edges(d["pos64"], d["faces"], d["node_ptr"], d["face_ptr"], periodic)
    n = d["pos64"].shape[0]
    b = d["node_ptr"].numel() - 1
    counts = d["node_ptr"][1:] - d["node_ptr"][:-1]
    labels = d["labels"].unsqueeze(1)
    op = None
    if with_op_div:
        op = torch.sparse_coo_tensor(torch.stack([d["op_row"], d["op_col"]]), d["op_val"], (n, int(h["op_width"])),
                                     is_coalesced=True)
    return MeshBatch(pos=d["pos64"].to(torch.float32), edge_index=edge_index, edge_attr=edge_attr,
                     mean_stress=d["mean_stress"], local_stress=d["local_stress"], nodes_types=labels,
                     surfaces_nodes_for_div=labels, op_div_matrix=op, ptr=d["node_ptr"],
                     batch=torch.repeat_interleave(torch.arange(b, device=device), counts, output_size=n),
                     batch_size=b, num_nodes=n, is_periodic=periodic)


class ResidentDataset:
    """The whole dataset in HBM (a 10 000-mesh set of ~1 000-node meshes is ~1 GB without / ~4 GB with the divergence
    operators, of 180 GB): every sample's coordinates, faces, fields, labels and operator triplets are uploaded ONCE,
    concatenated; a batch is then gathered on the device from the per-sample ranges by ONE launch
    (``pdg_resident_gather``) -- per step the host sends one small integer descriptor instead of collating, pinning and
    copying ~3 MB (12 MB with the operator), and the device batcher (edges, periodic edges, collation) runs exactly as
    for host batches.  Same MeshBatch, bit for bit."""

    def __init__(self, samples, device="cuda", with_op_div: bool = True, chunk: int = 512):
        self.device = torch.device(device)
        self.with_op = with_op_div
        self.n_nodes = np.array([np.asarray(s["pos"]).shape[0] for s in samples], dtype=np.int64)
        self.n_faces = np.array([np.asarray(s["faces"]).shape[1] for s in samples], dtype=np.int64)
        self.n_nnz = np.array([len(s["op_div_row"]) if with_op_div else 0 for s in samples], dtype=np.int64)
        self.node_start = np.concatenate([[0], np.cumsum(self.n_nodes)])[:-1]
        self.face_start = np.concatenate([[0], np.cumsum(self.n_faces)])[:-1]
        self.nnz_start = np.concatenate([[0], np.cumsum(self.n_nnz)])[:-1]
        parts = {k: [] for k in ("pos64", "faces", "mean_stress", "local_stress", "labels", "op_row", "op_col", "op_val")}
        for i in range(0, len(samples), chunk):  # bounded pinned staging
            h = host_arrays(samples[i:i + chunk], with_op_div)
            for k in parts:
                if k in h:
                    t = h[k].to(self.device, non_blocking=True)
                    if k == "op_row":
                        t = t + int(self.node_start[i])  # rows are dataset-global node ids
                    parts[k].append(t)
            torch.cuda.synchronize(self.device)  # the pinned chunk may be freed
        self.t = {k: torch.cat(v, dim=1 if k == "faces" else 0) for k, v in parts.items() if v}

    def nbytes(self) -> int:
        return int(sum(t.numel() * t.element_size() for t in self.t.values()))

    def batch(self, ids, periodic: bool = True, with_op_div: bool = True) -> MeshBatch:
        """One launch (``pdg_resident_gather``) + the device batcher; the host sends a 6 B + 3 integer descriptor."""
        if with_op_div and not self.with_op:
            raise ValueError("this ResidentDataset was built without the divergence operators")
        dev, L, t = self.device, _lib.lib(), self.t
        ids = np.asarray(ids, dtype=np.int64)
        b = int(ids.size)
        n_i, f_i = self.n_nodes[ids], self.n_faces[ids]
        z_i = self.n_nnz[ids] if with_op_div else np.zeros(b, dtype=np.int64)
        n, f, z = int(n_i.sum()), int(f_i.sum()), int(z_i.sum())
        ptr = lambda v: np.concatenate([[0], np.cumsum(v)])  # noqa: E731
        meta = np.concatenate([ptr(n_i), ptr(f_i), ptr(z_i), self.node_start[ids], self.face_start[ids],
                               self.nnz_start[ids]]).astype(np.int64)
        with torch.cuda.device(dev):
            m = torch.from_numpy(meta).to(dev, non_blocking=True)
            npf = int(t["faces"].shape[0])
            e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
            pos64, pos32 = e((n, 2), torch.float64), e((n, 2), torch.float32)
            ms, ls = e((n, 3), torch.float32), e((n, 3), torch.float32)
            labels, bvec, faces = e((n,), torch.int64), e((n,), torch.int64), e((npf, f), torch.int64)
            op_idx = e((2, z), torch.int64) if z else None
            op_val = e((z,), torch.float32) if z else None
            _lib.check(L.pdg_resident_gather(
                _lib.ptr(t["pos64"]), _lib.ptr(t["mean_stress"]), _lib.ptr(t["local_stress"]), _lib.ptr(t["labels"]),
                _lib.ptr(t["faces"]), int(t["faces"].shape[1]), npf, _lib.ptr(t.get("op_row")), _lib.ptr(t.get("op_col")),
                _lib.ptr(t.get("op_val")), _lib.ptr(m), b, n, f, z, _lib.ptr(pos64), _lib.ptr(pos32), _lib.ptr(ms), _lib.ptr(ls),
                _lib.ptr(labels), _lib.ptr(bvec), _lib.ptr(faces), _lib.ptr(op_idx), _lib.ptr(op_val), _lib.stream_ptr(dev)),
                "pdg_resident_gather")
        node_ptr, face_ptr = m[:b + 1], m[b + 1:2 * b + 2]
        edge_index, edge_attr = build_edges(pos64, faces, node_ptr, face_ptr, periodic)
        labels = labels.unsqueeze(1)
        op = None
        if with_op_div:
            op = torch.sparse_coo_tensor(op_idx, op_val, (n, int(2 * n_i.max())), is_coalesced=True)
        return MeshBatch(pos=pos32, edge_index=edge_index, edge_attr=edge_attr, mean_stress=ms, local_stress=ls,
                         nodes_types=labels, surfaces_nodes_for_div=labels, op_div_matrix=op, ptr=node_ptr, batch=bvec,
                         batch_size=b, num_nodes=n, is_periodic=periodic)


def dataset_stats(batches) -> dict:
    """The 8 scalar statistics of datasets.py:283-291 (mean / unbiased std over the whole set)."""
    cat = lambda k: torch.cat([getattr(b, k).reshape(-1) for b in batches])  # noqa: E731
    pos, ms, ls, ew = cat("pos"), cat("mean_stress"), cat("local_stress"), cat("edge_attr")
    return dict(mean_pos=pos.mean(), std_pos=pos.std(), mean_mean_stress=ms.mean(), std_mean_stress=ms.std(),
                mean_local_stress=ls.mean(), std_local_stress=ls.std(), mean_edge_weight=ew.mean(),
                std_edge_weight=ew.std())


class DevicePrefetcher:
    """Builds the next batches (pinned-host -> device copies, device edge construction, graph plan) on a side
    stream while step j trains on the main stream -- the GPU-side analogue of a DataLoader worker.

        pf = DevicePrefetcher(host_batches, device)
        for j in range(steps):
            batch = pf.get()         # ready on the current stream
            loss = train_step(batch) # enqueue the step first ...
            pf.prefetch()            # ... then top the staging queue up underneath it

    The staging runs in ONE worker thread, `depth` batches ahead: building a batch ends with a host read of the
    edge count (pdg_batch_count) that has to wait for the side stream's kernels to find free SMs between the
    training kernels; in the training thread that wait let the main stream's queue run dry now and then
    (6.99 vs 5.4 ms/step end to end, run to run).  threaded=False keeps everything in the calling thread.
    """

    def __init__(self, host_batches, device="cuda", periodic=True, with_op_div=True, build_plans=True, n_batches=None,
                 depth=2, threaded=True):
        self.host, self.device = host_batches, torch.device(device)
        if self.device.index is None:  # the worker thread needs an explicit ordinal (its current device starts at 0)
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.periodic, self.with_op, self.build_plans = periodic, with_op_div, build_plans
        self.n_batches = n_batches  # None: rotate over host_batches forever; else stop staging after that many
        self.stream = torch.cuda.Stream(self.device, priority=-1)  # high priority: its small kernels slot in at the main stream's kernel boundaries
        self.j = 0
        self.depth = max(1, int(depth))
        self._q = collections.deque()
        self._pool = concurrent.futures.ThreadPoolExecutor(1, thread_name_prefix="pdg-prefetch") if threaded else None
        self.prefetch()

    def _build(self, j):
        torch.cuda.set_device(self.device)
        with torch.cuda.stream(self.stream):
            h = self.host[j]  # host collation (or nothing, for a device-resident dataset) happens here, in the worker
            b = h() if callable(h) else batch_from_host(h, self.device, self.periodic, self.with_op)
            if self.build_plans:
                from .autograd import build_plan
                b._pdg_plan_buf = build_plan(b.edge_index, b.num_nodes).buf
                if self.with_op and b.op_div_matrix is not None:
                    from .loss import build_opdiv_plan
                    b._pdg_opplan_buf = build_opdiv_plan(b.op_div_matrix, b.ptr).buf
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return b, ev

    def prefetch(self):
        while len(self._q) < self.depth and (self.n_batches is None or self.j < self.n_batches):
            j = self.j % len(self.host)
            self.j += 1
            self._q.append(self._pool.submit(self._build, j) if self._pool is not None else self._build(j))

    def get(self) -> MeshBatch:
        self.prefetch()
        if not self._q:
            raise StopIteration("DevicePrefetcher: all n_batches batches were consumed")
        item = self._q.popleft()
        b, ev = item.result() if self._pool is not None else item
        main = torch.cuda.current_stream(self.device)
        main.wait_event(ev)
        for v in vars(b).values():  # the tensors were allocated on the side stream
            if torch.is_tensor(v):
                if v.is_sparse:
                    v._indices().record_stream(main)
                    v._values().record_stream(main)
                else:
                    v.record_stream(main)
        return b

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
