"""On-disk dataset format: legacy-VTK reader/writer, .npz fields, dataset.csv.
Reference: generate_dataset.py:558-598 (writer), datasets.py:240-281 (reader), convert_utils.py:26-60 (faces)."""
import os

import numpy as np
import pytest
import torch

import pdg_helpers as H
from oracle import pdg_oracle as O
from pdivgnn_b200 import io as pio
from pdivgnn_b200 import synth


@pytest.mark.parametrize("binary", [True, False])
@pytest.mark.parametrize("version", ["4.2", "5.1"])
@pytest.mark.parametrize("dataset", ["UNSTRUCTURED_GRID", "POLYDATA"])
@pytest.mark.parametrize("quads", [False, True])
def test_vtk_roundtrip_is_bit_exact(tmp_path, binary, version, dataset, quads):
    s = synth.make_quad_rve_mesh(5, 200) if quads else synth.make_rve_mesh(5, 200)
    f = str(tmp_path / "m.vtk")
    pio.write_legacy_vtk(f, s["pos"], s["faces"], binary=binary, version=version, dataset=dataset)
    pts, faces = pio.read_legacy_vtk(f)
    assert pts.dtype == np.float64 and faces.dtype == np.int64 and faces.shape[0] == (4 if quads else 3)
    assert np.array_equal(pts, np.asarray(s["pos"], dtype=np.float64))  # %.17g / raw doubles: exact
    assert np.array_equal(faces, np.asarray(s["faces"]))


def test_known_answer_ascii_file(tmp_path):
    """A hand-written 2-triangle square in the classic ASCII layout (with attribute sections to skip)."""
    f = tmp_path / "sq.vtk"
    f.write_text("# vtk DataFile Version 3.0\nsquare\nASCII\nDATASET POLYDATA\nPOINTS 4 float\n0 0 0 1 0 0\n1 1 0\n0 1 0\n"
                 "POLYGONS 2 8\n3 0 1 2\n3 0 2 3\nPOINT_DATA 4\nSCALARS u float 1\nLOOKUP_TABLE default\n1 2 3 4\n")
    pts, faces = pio.read_legacy_vtk(str(f))
    assert pts.tolist() == [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]]
    assert faces.tolist() == [[0, 0], [1, 2], [2, 3]]
    # graph of that mesh through the oracle: 5 undirected edges -> 10 directed
    ei = O.face_to_edge(torch.from_numpy(faces), 4)
    assert ei.shape == (2, 10)


def test_quad_file_known_answer(tmp_path):
    """One quad as POLYDATA: faces [4,1]; its graph has the 4 sides only (convert_utils.py:62-81), no diagonal."""
    q = tmp_path / "quad.vtk"
    q.write_text("# vtk DataFile Version 3.0\nq\nASCII\nDATASET POLYDATA\nPOINTS 4 float\n0 0 0 1 0 0 1 1 0 0 1 0\nPOLYGONS 1 5\n4 0 1 2 3\n")
    pts, faces = pio.read_legacy_vtk(str(q))
    assert faces.tolist() == [[0], [1], [2], [3]]
    ei = O.quad_face_to_edge(torch.from_numpy(faces), 4)
    assert ei.t().tolist() == [[0, 1], [0, 3], [1, 0], [1, 2], [2, 1], [2, 3], [3, 0], [3, 2]]


def test_rejects_what_the_hot_path_cannot_use(tmp_path):
    q = tmp_path / "mixed.vtk"  # a triangle and a quad: the reference handles single element-type meshes only
    q.write_text("# vtk DataFile Version 3.0\nq\nASCII\nDATASET POLYDATA\nPOINTS 5 float\n0 0 0 1 0 0 1 1 0 0 1 0 2 0 0\n"
                 "POLYGONS 2 9\n4 0 1 2 3\n3 1 4 2\n")
    with pytest.raises(NotImplementedError):
        pio.read_legacy_vtk(str(q))
    q = tmp_path / "penta.vtk"
    q.write_text("# vtk DataFile Version 3.0\nq\nASCII\nDATASET POLYDATA\nPOINTS 5 float\n0 0 0 1 0 0 1 1 0 0 1 0 2 0 0\n"
                 "POLYGONS 1 6\n5 0 1 4 2 3\n")
    with pytest.raises(NotImplementedError):
        pio.read_legacy_vtk(str(q))
    q = tmp_path / "wrongtype.vtk"  # 4-point cells declared as VTK_TETRA (10): not a quad mesh
    q.write_text("# vtk DataFile Version 3.0\nq\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS 4 float\n0 0 0 1 0 0 1 1 0 0 1 0\n"
                 "CELLS 1 5\n4 0 1 2 3\nCELL_TYPES 1\n10\n")
    with pytest.raises(NotImplementedError):
        pio.read_legacy_vtk(str(q))
    bad = tmp_path / "bad.vtk"
    bad.write_text("hello\n")
    with pytest.raises(ValueError):
        pio.read_legacy_vtk(str(bad))
    s = synth.make_rve_mesh(1, 120)
    f = str(tmp_path / "t.vtk")
    pio.write_legacy_vtk(f, s["pos"], s["faces"], binary=True)
    data = open(f, "rb").read()
    open(f, "wb").write(data[:len(data) // 2])
    with pytest.raises(ValueError):
        pio.read_legacy_vtk(f)
    g = tmp_path / "grid.vtk"
    g.write_text("# vtk DataFile Version 3.0\ng\nASCII\nDATASET STRUCTURED_POINTS\nDIMENSIONS 2 2 1\n")
    with pytest.raises(NotImplementedError):
        pio.read_legacy_vtk(str(g))


def test_sample_and_dataset_roundtrip_gives_the_same_graph(tmp_path):
    samples = synth.make_dataset(3, 150, 11)
    csv = pio.write_dataset(samples, str(tmp_path / "ds"), binary=True, version="5.1")
    import pandas as pd
    df = pd.read_csv(csv)
    assert list(df.columns[:2]) == ["mesh_filename", "data_filename"] and len(df) == 3
    back = [pio.read_sample(m, d) for m, d in zip(df["mesh_filename"], df["data_filename"])]
    for a, b in zip(samples, back):
        ga, gb = O.build_graph(a, True), O.build_graph(b, True)
        assert torch.equal(ga.edge_index, gb.edge_index) and torch.equal(ga.edge_attr, gb.edge_attr)
        assert torch.equal(ga.local_stress, gb.local_stress) and torch.equal(ga.mean_stress, gb.mean_stress)
        assert torch.equal(ga.nodes_types, gb.nodes_types)
        assert torch.equal(ga.op_div_matrix.indices(), gb.op_div_matrix.indices())
        assert torch.equal(ga.op_div_matrix.values(), gb.op_div_matrix.values())
    sa, sb = O.dataset_stats([O.build_graph(s, True) for s in samples]), O.dataset_stats([O.build_graph(s, True) for s in back])
    assert all(torch.equal(sa[k], sb[k]) for k in sa)


def test_sample_validation(tmp_path):
    s = synth.make_rve_mesh(2, 120)
    mf, df = str(tmp_path / "a.vtk"), str(tmp_path / "a.npz")
    pio.write_sample(s, mf, df)
    np.savez(df, stress_field=np.zeros((3, 3)), mean_stress=np.zeros(3))
    with pytest.raises(KeyError):
        pio.read_sample(mf, df)
    s2 = dict(s)
    s2["stress_field"] = np.asarray(s["stress_field"])[:-1]
    pio.write_sample(s2, str(tmp_path / "b.vtk"), str(tmp_path / "b.npz"))
    with pytest.raises(ValueError):
        pio.read_sample(str(tmp_path / "b.vtk"), str(tmp_path / "b.npz"))


@pytest.mark.parametrize("threaded", [True, False])
def test_device_prefetcher_queue_logic(monkeypatch, threaded):
    """Host-side logic of batcher.DevicePrefetcher (staging depth, order, rotation, exhaustion) with the CUDA pieces
    stubbed out: the batches come back in order, never more than `depth` are staged ahead, n_batches is respected."""
    import threading
    from pdivgnn_b200 import batcher

    class _Stream:
        def __init__(self, *a, **k):
            pass

        def wait_event(self, ev):
            ev.waited = True

    class _Ev:
        waited = False

    built, lock = [], threading.Lock()

    def fake_build(self, j):
        h = self.host[j]
        with lock:
            built.append(h)
        return batcher.MeshBatch(tag=h, batch_size=1), _Ev()

    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: _Stream())
    monkeypatch.setattr(batcher.DevicePrefetcher, "_build", fake_build)
    hosts = ["h0", "h1", "h2"]
    pf = batcher.DevicePrefetcher(hosts, torch.device("cuda", 0), n_batches=7, depth=2, threaded=threaded)
    got = []
    for j in range(7):
        b = pf.get()
        got.append(b.tag)
        with lock:
            assert len(built) <= j + 1 + 2  # at most `depth` batches staged beyond the one handed out
        pf.prefetch()
    assert got == ["h0", "h1", "h2", "h0", "h1", "h2", "h0"]  # rotation over the host batches, in order
    pf.prefetch()  # no-op once n_batches were staged
    with pytest.raises(StopIteration):
        pf.get()
    pf.close()
    assert len(built) == 7
    # endless rotation when n_batches is None; depth is clamped to >= 1
    pf2 = batcher.DevicePrefetcher(hosts, torch.device("cuda", 0), depth=0, threaded=threaded)
    assert [pf2.get().tag for _ in range(5)] == ["h0", "h1", "h2", "h0", "h1"]
    pf2.close()


class _FakeDataset:
    """len-only stand-in: DeviceLoader's sharding arithmetic needs nothing else (no CUDA)."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


@pytest.mark.parametrize("n_items,batch,world", [(10 * 4, 4, 8), (37, 4, 8), (33, 32, 2), (5, 1, 4), (64, 8, 8), (3, 2, 4)])
@pytest.mark.parametrize("uneven", ["pad", "drop"])
def test_loader_gives_every_rank_the_same_number_of_steps(n_items, batch, world, uneven):
    """the gradient all-reduce runs inside every backward, so ranks must not disagree on the step count
    (10 batches on 8 GPUs used to leave 6 ranks waiting in NCCL)."""
    from pdivgnn_b200.io import DeviceLoader
    loaders = [DeviceLoader(_FakeDataset(n_items), batch, True, 69, False, r, world, False, uneven) for r in range(world)]
    lens = [len(l) for l in loaders]
    sched = [l._batches() for l in loaders]
    assert len(set(lens)) == 1 and all(len(s) == lens[0] for s in sched)
    nb = (n_items + batch - 1) // batch
    assert lens[0] == (nb // world if uneven == "drop" else (nb + world - 1) // world)
    seen = sorted(int(i) for s in sched for c in s for i in c)
    if uneven == "pad":
        assert set(seen) == set(range(n_items))          # every sample is visited, a few twice
        assert len(seen) - n_items < world * batch
    else:
        assert len(seen) == len(set(seen))                # nothing twice, at most world-1 batches dropped
        assert n_items - len(seen) < world * batch
    # the same epoch order on every rank (same seed): ranks see disjoint batches within a step
    for s in range(lens[0]):
        firsts = [int(sched[r][s][0]) for r in range(world)]
        assert len(set(firsts)) == world or nb < world


def test_loader_rejects_bad_rank_and_policy():
    from pdivgnn_b200.io import DeviceLoader
    with pytest.raises(ValueError):
        DeviceLoader(_FakeDataset(8), 2, False, 0, False, 2, 2, False)
    with pytest.raises(ValueError):
        DeviceLoader(_FakeDataset(8), 2, False, 0, False, 0, 2, False, "wrap")


_FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vtk")
_FIX_POINTS = [[0, 0, 0], [1.5, 0, 0], [3, 0, 0], [0, 1, 0], [1.5, 1, 0], [3, 1, 0]]
_FIX_TRIS = [[0, 0, 1, 1], [1, 4, 2, 5], [4, 3, 5, 4]]
_FIX_QUADS = [[0, 1], [1, 2], [4, 5], [3, 4]]


@pytest.mark.parametrize("name,faces", [
    ("tri_ug_v51_ascii.vtk", _FIX_TRIS), ("tri_ug_v51_binary.vtk", _FIX_TRIS), ("tri_ug_v42_ascii.vtk", _FIX_TRIS),
    ("tri_pd_v51_ascii.vtk", _FIX_TRIS), ("quad_ug_v42_binary.vtk", _FIX_QUADS), ("quad_ug_v51_ascii.vtk", _FIX_QUADS),
    ("quad_pd_v30_ascii.vtk", _FIX_QUADS)])
def test_committed_vtk_fixtures_in_the_vtk_writer_layout(name, faces):
    """Files written byte by byte in the layouts vtkDataWriter emits (tests/golden/make_vtk_fixtures.py, independent of
    io.write_legacy_vtk): 5.1 OFFSETS/CONNECTIVITY and classic cell streams, ASCII lines with trailing blanks, big-endian
    BINARY, METADATA blocks, CELL_TYPES one per line, trailing POINT_DATA.  Expected arrays are written out by hand."""
    pts, f = pio.read_legacy_vtk(os.path.join(_FIX, name))
    assert pts.tolist() == _FIX_POINTS and f.tolist() == faces and f.dtype == np.int64
    # and the graph the reference builds from it: the 3 x 2 patch has 7 sides (+ 2 diagonals when triangulated)
    ei = (O.face_to_edge if len(faces) == 3 else O.quad_face_to_edge)(torch.from_numpy(f), 6)
    assert ei.shape[1] == (18 if len(faces) == 3 else 14)


def test_node_labels_two_holes_and_a_hole_touching_no_side():
    """Hand-derived labels (datasets.py:133-179 semantics): an 8 x 8-node plate of quads with two single-cell holes.
    Region touching the bounding box -> 1 on the 28 side nodes; the 4 corners of each removed cell -> -1; the reference
    itself would stop at ``assert n_regions == 2`` for this mesh (3 loops) -- the restatement reports 3 regions."""
    pos, quads, tris, want = H.two_hole_plate()
    assert (want == 1).sum() == 28 and (want == -1).sum() == 8
    for faces in (quads, tris):  # the same plate triangulated has the same boundary loops
        labels, nreg = O.compute_node_labels(pos, faces)
        assert nreg == 3 and np.array_equal(labels, want)
    # one hole only, far from every side: exactly the reference's 2-region case
    pos, quads, tris, want = H.two_hole_plate(holes=((3, 3),))
    labels, nreg = O.compute_node_labels(pos, quads)
    assert nreg == 2 and np.array_equal(labels, want)
    assert sorted(np.nonzero(labels == -1)[0].tolist()) == [27, 28, 35, 36]
