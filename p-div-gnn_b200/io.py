"""On-disk dataset format of the reference -> host arrays for the device batcher.

The reference stores one sample as a pair (generate_dataset.py:558-598):

  ``<name>.vtk``  the triangle (or quad) mesh written by ``pyvista.DataSet.save`` (legacy VTK file, ASCII or BINARY,
                  ``POLYDATA`` or ``UNSTRUCTURED_GRID``, file versions <= 4.2 and 5.1), read back with
                  ``pv.get_reader(...).read()`` (datasets.py:252) and turned into ``faces [3,F]`` / ``[4,F]`` by
                  ``convert_utils._format_faces_from_pyvista`` / ``mesh_to_graph`` (convert_utils.py:26-60);
  ``<name>.npz``  ``stress_field [N,3]``, ``mean_stress [3]``, ``op_div_matrix_{data,row_indices,col_indices,
                  shape}``, ``node_labels [N]`` (+ ``mean_strain``, ``mean_stress_material``, ``op_mean_stress``).

and lists the pairs in ``dataset.csv`` (columns ``mesh_filename``, ``data_filename`` + generation parameters,
generate_dataset.py:48-70).  pyvista / vtk are not needed here: :func:`read_legacy_vtk` is a small parser of the
legacy format.  :class:`MeshStressFieldDataset` mirrors ``MeshStressFieldDatasetInMemory`` (datasets.py:233-298):
same constructor arguments, the same 8 statistics attributes, ``len`` / indexing, ``dataframe``; its ``loader``
yields batched graphs built ON THE GPU by :mod:`batcher` (the PyG ``DataLoader`` + collate replacement).
:func:`write_legacy_vtk` / :func:`write_sample` are the matching writers (tests, synthetic datasets).
"""
from __future__ import annotations

import os
import re
from typing import Iterable, Sequence

import numpy as np

_VTK_TRIANGLE, _VTK_QUAD = 5, 9  # VTK cell type ids (pv.CellType.TRIANGLE / QUAD; mesh_to_graph dispatches on QUAD, convert_utils.py:52)
_DTYPES = {"bit": None, "unsigned_char": "u1", "char": "i1", "unsigned_short": "u2", "short": "i2", "unsigned_int": "u4",
           "int": "i4", "unsigned_long": "u8", "long": "i8", "float": "f4", "double": "f8", "vtktypeint64": "i8",
           "vtktypeint32": "i4", "vtktypeuint8": "u1"}


class _Cursor:
    """Token / raw-byte reader over a legacy VTK file (BINARY payloads are big-endian)."""

    def __init__(self, data: bytes):
        self.d, self.i = data, 0

    def eof(self):
        return self.i >= len(self.d)

    def line(self) -> str:
        j = self.d.find(b"\n", self.i)
        j = len(self.d) if j < 0 else j
        s = self.d[self.i:j].decode("ascii", "replace").strip()
        self.i = j + 1
        return s

    def next_nonempty(self) -> str:
        while not self.eof():
            s = self.line()
            if s:
                return s
        return ""

    def array(self, n: int, dtype: str, binary: bool) -> np.ndarray:
        code = _DTYPES.get(dtype.lower())
        if code is None:
            raise NotImplementedError(f"legacy VTK data type '{dtype}'")
        if binary:
            nb = n * np.dtype(code).itemsize
            if self.i + nb > len(self.d):
                raise ValueError("truncated BINARY section in VTK file")
            a = np.frombuffer(self.d, dtype=">" + code, count=n, offset=self.i).astype(code)
            self.i += nb
            if not self.eof() and self.d[self.i:self.i + 1] == b"\n":
                self.i += 1
            return a
        vals = []
        while len(vals) < n:
            if self.eof():
                raise ValueError("truncated ASCII section in VTK file")
            vals.extend(self.line().split())
        if len(vals) != n:
            raise ValueError("ASCII section longer than declared")
        return np.array(vals, dtype=np.float64).astype(code) if code[0] == "f" else np.array([int(float(v)) for v in vals],
                                                                                             dtype=code)


def _cells_to_faces(conn: np.ndarray, offsets: np.ndarray | None, n_cells: int, types: np.ndarray | None) -> np.ndarray:
    """Cell connectivity -> ``[k,F]`` int64 with k = 3 (all triangles) or 4 (all quads): the reference handles
    "single element-type meshes" only (convert_utils.py:24) and reshapes the stream by the first cell's size
    (``_format_faces_from_pyvista``, :27-33).  ``offsets is None``: classic ``n i j k n i j k ...`` stream."""
    if n_cells <= 0:
        raise ValueError("VTK file without cells")
    if offsets is None:
        k = int(conn[0])
        if conn.size != n_cells * (k + 1):
            raise NotImplementedError("mixed cell sizes: only all-triangle or all-quad meshes are supported (convert_utils.py:24)")
        tab = conn.reshape(n_cells, k + 1)
        if np.any(tab[:, 0] != k):
            raise NotImplementedError("mixed cell sizes: only all-triangle or all-quad meshes are supported (convert_utils.py:24)")
        t = tab[:, 1:]
    else:
        sizes = np.diff(offsets)
        k = int(sizes[0])
        if np.any(sizes != k) or sizes.size != n_cells:
            raise NotImplementedError("mixed cell sizes: only all-triangle or all-quad meshes are supported (convert_utils.py:24)")
        t = conn.reshape(n_cells, k)
    if k not in (3, 4):
        raise NotImplementedError(f"cells with {k} points: only triangles and quads are supported (convert_utils.py:47-81)")
    if types is not None:
        want = _VTK_TRIANGLE if k == 3 else _VTK_QUAD
        if types.size != n_cells or np.any(types != want):
            raise NotImplementedError(f"cell types {sorted(set(types.tolist()))} with {k}-point cells: expected VTK type {want} only")
    return np.ascontiguousarray(t.astype(np.int64).T)


def read_legacy_vtk(path: str):
    """(points [N,3] float64, faces [3,F] or [4,F] int64) of a legacy ``.vtk`` triangle or quad mesh.

    Same result as ``pv.get_reader(path).read()`` followed by ``_format_faces_from_pyvista`` (convert_utils.py:26-44)
    for an all-triangle or all-quad mesh; point / cell data sections are skipped.
    """
    with open(path, "rb") as f:
        cur = _Cursor(f.read())
    head = cur.line()
    m = re.match(r"#\s*vtk\s+DataFile\s+Version\s+(\d+)\.(\d+)", head)
    if not m:
        raise ValueError(f"{path}: not a legacy VTK file")
    cur.line()  # title
    fmt = cur.next_nonempty().upper()
    if fmt not in ("ASCII", "BINARY"):
        raise ValueError(f"{path}: unknown encoding '{fmt}'")
    binary = fmt == "BINARY"
    ds = cur.next_nonempty().split()
    if len(ds) != 2 or ds[0].upper() != "DATASET" or ds[1].upper() not in ("POLYDATA", "UNSTRUCTURED_GRID"):
        raise NotImplementedError(f"{path}: dataset '{' '.join(ds)}' (expected POLYDATA or UNSTRUCTURED_GRID)")
    points, faces, types, pending = None, None, None, None
    while not cur.eof():
        s = cur.next_nonempty()
        if not s:
            break
        tok = s.split()
        key = tok[0].upper()
        if key == "POINTS":
            n = int(tok[1])
            points = cur.array(3 * n, tok[2], binary).astype(np.float64).reshape(n, 3)
        elif key in ("POLYGONS", "CELLS", "TRIANGLE_STRIPS", "LINES", "VERTICES"):
            a, b = int(tok[1]), int(tok[2])
            nxt_pos = cur.i
            nxt = cur.next_nonempty().split()
            if nxt and nxt[0].upper() == "OFFSETS":  # file version 5.x: CELLS <n_offsets> <n_conn> / OFFSETS / CONNECTIVITY
                off = cur.array(a, nxt[1], binary).astype(np.int64)
                cn = cur.next_nonempty().split()
                if not cn or cn[0].upper() != "CONNECTIVITY":
                    raise ValueError(f"{path}: CONNECTIVITY expected after OFFSETS")
                conn = cur.array(b, cn[1], binary).astype(np.int64)
                if key in ("POLYGONS", "CELLS"):
                    pending = (conn, off, a - 1)
            else:  # classic: <n_cells> <size> then the int stream
                cur.i = nxt_pos
                conn = cur.array(b, "int", binary).astype(np.int64)
                if key in ("POLYGONS", "CELLS"):
                    pending = (conn, None, a)
                elif a:
                    raise NotImplementedError(f"{path}: non-empty {key} section")
        elif key == "CELL_TYPES":
            types = cur.array(int(tok[1]), "int", binary)
        elif key == "METADATA":  # information keys of the preceding array: skipped up to the blank line that ends them
            while not cur.eof() and cur.line():
                pass
        elif key in ("POINT_DATA", "CELL_DATA", "FIELD"):
            break  # attributes follow the geometry; nothing else is needed
        else:
            raise NotImplementedError(f"{path}: unsupported section '{key}'")
    if points is None or pending is None:
        raise ValueError(f"{path}: POINTS and POLYGONS/CELLS are required")
    conn, off, ncell = pending
    faces = _cells_to_faces(conn, off, ncell, types)
    if faces.size and (faces.min() < 0 or faces.max() >= points.shape[0]):
        raise ValueError(f"{path}: face index out of range")
    return points, faces


def write_legacy_vtk(path: str, points: np.ndarray, faces: np.ndarray, binary: bool = True, version: str = "4.2",
                     dataset: str = "UNSTRUCTURED_GRID") -> None:
    """Write a triangle ([3,F]) or quad ([4,F]) mesh the way ``pyvista`` / VTK does (``version`` '4.2' classic or
    '5.1' offsets layout)."""
    pts = np.asarray(points, dtype=np.float64)
    if pts.shape[1] == 2:
        pts = np.concatenate([pts, np.zeros((pts.shape[0], 1))], axis=1)
    faces = np.asarray(faces, dtype=np.int64)
    k = faces.shape[0]
    if k not in (3, 4):
        raise NotImplementedError("faces must be [3,F] or [4,F]")
    tri = faces.T.reshape(-1, k)
    n, f = pts.shape[0], tri.shape[0]

    def emit(fh, arr, code):
        if binary:
            fh.write(np.ascontiguousarray(arr).astype(">" + code).tobytes())
            fh.write(b"\n")
        else:
            a2 = np.asarray(arr).reshape(-1)
            fmt = "%.17g" if code[0] == "f" else "%d"
            for k in range(0, a2.size, 9):
                fh.write((" ".join(fmt % v for v in a2[k:k + 9]) + "\n").encode())

    with open(path, "wb") as fh:
        fh.write(f"# vtk DataFile Version {version}\nvtk output\n{'BINARY' if binary else 'ASCII'}\nDATASET {dataset}\n".encode())
        fh.write(f"POINTS {n} double\n".encode())
        emit(fh, pts, "f8")
        sec = "POLYGONS" if dataset == "POLYDATA" else "CELLS"
        if version.startswith("5"):
            fh.write(f"{sec} {f + 1} {k * f}\nOFFSETS vtktypeint64\n".encode())
            emit(fh, np.arange(0, k * f + 1, k), "i8")
            fh.write(b"CONNECTIVITY vtktypeint64\n")
            emit(fh, tri, "i8")
        else:
            fh.write(f"{sec} {f} {(k + 1) * f}\n".encode())
            emit(fh, np.concatenate([np.full((f, 1), k, np.int64), tri], axis=1), "i4")
        if dataset == "UNSTRUCTURED_GRID":
            fh.write(f"CELL_TYPES {f}\n".encode())
            emit(fh, np.full(f, _VTK_TRIANGLE if k == 3 else _VTK_QUAD), "i4")


_NPZ_KEYS = ("stress_field", "mean_stress", "op_div_matrix_data", "op_div_matrix_col_indices", "op_div_matrix_row_indices",
             "op_div_matrix_shape", "node_labels")


def read_sample(mesh_filename: str, data_filename: str) -> dict:
    """One dataset row -> the sample dict :func:`batcher.host_arrays` consumes (same keys as :mod:`synth`)."""
    pos, faces = read_legacy_vtk(mesh_filename)
    with np.load(data_filename) as z:
        missing = [k for k in _NPZ_KEYS if k not in z.files]
        if missing:
            raise KeyError(f"{data_filename}: missing arrays {missing}")
        d = {k: z[k] for k in _NPZ_KEYS}
    n = pos.shape[0]
    if d["stress_field"].shape != (n, 3) or d["node_labels"].shape[0] != n:
        raise ValueError(f"{data_filename}: fields do not match the {n} mesh points of {mesh_filename}")
    return dict(pos=pos, faces=faces, labels=d["node_labels"].astype(np.int64),
                op_div_row=d["op_div_matrix_row_indices"].astype(np.int64),
                op_div_col=d["op_div_matrix_col_indices"].astype(np.int64),
                op_div_data=d["op_div_matrix_data"].astype(np.float32),
                op_div_shape=np.asarray(d["op_div_matrix_shape"]).astype(np.int64),
                mean_stress=np.asarray(d["mean_stress"], dtype=np.float64), stress_field=d["stress_field"].astype(np.float32))


def write_sample(sample: dict, mesh_filename: str, data_filename: str, binary: bool = True, version: str = "4.2") -> None:
    """Counterpart of generate_dataset.py:584-598 for a sample dict (synthetic datasets, tests)."""
    write_legacy_vtk(mesh_filename, sample["pos"], sample["faces"], binary=binary, version=version)
    n = np.asarray(sample["pos"]).shape[0]
    np.savez(data_filename, stress_field=np.asarray(sample["stress_field"]), mean_stress=np.asarray(sample["mean_stress"]),
             op_div_matrix_data=np.asarray(sample["op_div_data"]), op_div_matrix_col_indices=np.asarray(sample["op_div_col"]),
             op_div_matrix_row_indices=np.asarray(sample["op_div_row"]),
             op_div_matrix_shape=np.asarray(sample.get("op_div_shape", (n, 2 * n))), node_labels=np.asarray(sample["labels"]))


def write_dataset(samples: Sequence[dict], folder: str, binary: bool = True, version: str = "4.2") -> str:
    """Write ``meshes/*.vtk``, ``fields/*.npz`` and ``dataset.csv`` (generate_dataset.py layout); returns the csv path."""
    import pandas as pd
    os.makedirs(os.path.join(folder, "meshes"), exist_ok=True)
    os.makedirs(os.path.join(folder, "fields"), exist_ok=True)
    rows = []
    for i, s in enumerate(samples):
        mf = os.path.join(folder, "meshes", f"hole_plate_mesh_{i}.vtk")
        df = os.path.join(folder, "fields", f"hole_plate_mesh_{i}.npz")
        write_sample(s, mf, df, binary=binary, version=version)
        ms = np.asarray(s["mean_stress"], dtype=np.float64)
        rows.append(dict(mesh_filename=mf, data_filename=df, mean_stress_x=ms[0], mean_stress_y=ms[1], mean_stress_xy=ms[2],
                         n_nodes=int(np.asarray(s["pos"]).shape[0]), n_elements=int(np.asarray(s["faces"]).shape[1])))
    csv = os.path.join(folder, "dataset.csv")
    pd.DataFrame(rows).to_csv(csv, index=False)
    return csv


class MeshStressFieldDataset:
    """``MeshStressFieldDatasetInMemory`` (datasets.py:233-298) without PyG / pyvista.

    ``dataframe``: a pandas DataFrame (or a csv path) with ``mesh_filename`` / ``data_filename`` columns.  Samples are
    read to host arrays once; graphs are built on the GPU batch by batch.  The 8 statistics are computed exactly as
    the reference does (mean / unbiased std over the concatenation of all graphs, edge weights incl. periodic
    zeros) -- on the device, through the same batcher that feeds training.
    """

    def __init__(self, dataframe, transform=None, periodic_graph: bool = True, device="cuda"):
        import pandas as pd
        import torch
        from . import batcher
        if transform is not None:
            raise NotImplementedError("transforms are not part of the hot path")
        if isinstance(dataframe, (str, os.PathLike)):
            dataframe = pd.read_csv(dataframe)
        self.dataframe = dataframe
        self.periodic_graph = periodic_graph
        self.device = torch.device(device)
        self.samples = [read_sample(m, d) for m, d in zip(dataframe["mesh_filename"], dataframe["data_filename"])]
        if not self.samples:
            raise ValueError("empty dataset")
        # statistics over the whole set (datasets.py:283-291): one pass of the device batcher in chunks
        acc = []
        for i in range(0, len(self.samples), 256):
            acc.append(batcher.batch_from_host(batcher.host_arrays(self.samples[i:i + 256]), self.device, periodic_graph, False))
        for k, v in batcher.dataset_stats(acc).items():
            setattr(self, k, v)
        del acc

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        return self.samples[i]

    def stats(self) -> dict:
        return {k: getattr(self, k) for k in ("mean_pos", "std_pos", "mean_mean_stress", "std_mean_stress",
                                              "mean_local_stress", "std_local_stress", "mean_edge_weight", "std_edge_weight")}

    def loader(self, batch_size: int, shuffle: bool = False, seed: int = 69, with_op_div: bool = True,
               rank: int = 0, world: int = 1, prefetch: bool = True, uneven: str = "pad",
               resident: bool | None = None) -> "DeviceLoader":
        """``resident`` (default: when the set fits in a quarter of the GPU's memory): keep the whole dataset in HBM and
        gather every batch there (:class:`batcher.ResidentDataset`) instead of collating and copying it from the host."""
        return DeviceLoader(self, batch_size, shuffle, seed, with_op_div, rank, world, prefetch, uneven, resident)

    def resident_store(self, with_op_div: bool):
        """The device-resident copy of the samples (built once; rebuilt with the operators when they are first needed)."""
        from . import batcher
        st = getattr(self, "_resident", None)
        if st is None or (with_op_div and not st.with_op):
            st = batcher.ResidentDataset(self.samples, self.device, with_op_div)
            self._resident = st
        return st

    def host_nbytes(self, with_op_div: bool) -> int:
        keys = ("pos", "faces", "stress_field", "labels") + (("op_div_row", "op_div_col", "op_div_data") if with_op_div else ())
        return int(sum(np.asarray(s[k]).nbytes for s in self.samples for k in keys))


class DeviceLoader:
    """``PyG.loader.DataLoader(dataset, batch_size, shuffle)`` replacement (gnn_train.py:387-394, gnn_inference.py:108-112).

    Iterating yields :class:`batcher.MeshBatch` objects on the GPU.  With ``prefetch`` the next batch (pinned-host ->
    device copies, device edge construction, graph plan) is staged on a side stream while the caller trains on
    the current one.  ``rank`` / ``world``: data-parallel sharding, batch j belongs to rank j % world.

    Every rank runs the SAME number of steps -- the gradient all-reduce sits inside the backward of every step, so a
    rank with one batch fewer would leave the others waiting in NCCL at the end of the epoch.  When the number of
    batches is not a multiple of ``world``: ``uneven="pad"`` (default, ``DistributedSampler`` semantics) wraps around
    and repeats batches from the start of the epoch's order, ``uneven="drop"`` drops the remainder (``drop_last``).
    """

    def __init__(self, dataset: MeshStressFieldDataset, batch_size: int, shuffle: bool, seed: int, with_op_div: bool,
                 rank: int, world: int, prefetch: bool, uneven: str = "pad", resident: bool | None = None):
        if uneven not in ("pad", "drop"):
            raise ValueError("uneven must be 'pad' or 'drop'")
        if world < 1 or not 0 <= rank < world:
            raise ValueError(f"invalid rank {rank} / world {world}")
        self.ds, self.batch_size, self.shuffle, self.seed = dataset, int(batch_size), shuffle, seed
        self.with_op, self.rank, self.world, self.prefetch, self.uneven = with_op_div, rank, world, prefetch, uneven
        self.epoch = 0
        self._host_cache = {}
        self._resident, self._store = resident, None

    @property
    def store(self):
        """The device-resident copy of the dataset this loader gathers its batches from, or None (host collation).
        Decided and built on first use: by default the set stays in HBM when it fits in a quarter of the GPU's memory."""
        if self._resident is None:
            import torch
            dev = torch.device(self.ds.device)
            self._resident = (dev.type == "cuda" and
                              self.ds.host_nbytes(self.with_op) < torch.cuda.get_device_properties(dev).total_memory // 4)
        if self._resident and self._store is None:
            self._store = self.ds.resident_store(self.with_op)
        return self._store

    def _steps_per_rank(self) -> int:
        nb = (len(self.ds) + self.batch_size - 1) // self.batch_size
        return nb // self.world if self.uneven == "drop" else (nb + self.world - 1) // self.world

    def _batches(self) -> list:
        idx = np.arange(len(self.ds))
        if self.shuffle:
            np.random.default_rng(self.seed + self.epoch).shuffle(idx)
        chunks = [idx[i:i + self.batch_size] for i in range(0, len(idx), self.batch_size)]
        steps = self._steps_per_rank()
        # global batch j of step s on rank r is chunks[(s * world + r) % len(chunks)]: identical step counts everywhere
        return [chunks[(s * self.world + self.rank) % len(chunks)] for s in range(steps)]

    def __len__(self):
        return self._steps_per_rank()

    def _host(self, chunk):
        from . import batcher
        key = tuple(int(i) for i in chunk)
        store = self.store
        if store is not None:  # device-resident dataset: the batch is gathered in HBM
            return lambda: store.batch(key, self.ds.periodic_graph, self.with_op)
        if self.shuffle:  # a fresh collation every epoch: pageable arrays (pinning costs more than the copy it saves)
            return batcher.host_arrays([self.ds.samples[i] for i in key], self.with_op, pin=False)
        if key not in self._host_cache:  # fixed order: pin every batch once
            self._host_cache[key] = batcher.host_arrays([self.ds.samples[i] for i in key], self.with_op)
        return self._host_cache[key]

    def __iter__(self) -> Iterable:
        from . import batcher
        chunks = self._batches()
        self.epoch += 1
        if not self.prefetch:
            for c in chunks:
                h = self._host(c)
                b = h() if callable(h) else batcher.batch_from_host(h, self.ds.device, self.ds.periodic_graph, self.with_op)
                b.sample_ids = [int(i) for i in c]
                yield b
            return
        hosts = _LazyHosts(self, chunks)
        pf = batcher.DevicePrefetcher(hosts, self.ds.device, self.ds.periodic_graph, self.with_op, n_batches=len(chunks))
        try:
            for c in chunks:
                b = pf.get()
                b.sample_ids = [int(i) for i in c]
                yield b            # the caller enqueues its step on this batch ...
                pf.prefetch()      # ... and the next batches are staged underneath it (no-op after the last one)
        finally:
            pf.close()


class _LazyHosts:
    def __init__(self, loader, chunks):
        self.loader, self.chunks = loader, chunks

    def __len__(self):
        return len(self.chunks)

    def __getitem__(self, j):
        return self.loader._host(self.chunks[j])
