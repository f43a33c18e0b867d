"""Dataset reader -> device loader -> train / evaluate / predict loops.
Reference: datasets.py:233-298, gnn_train.py:152-253, gnn_inference.py:45-81."""
import numpy as np
import pytest
import torch

import pdg_helpers as H
from oracle import pdg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def disk_dataset(tmp_path_factory):
    from pdivgnn_b200 import io as pio, synth
    samples = synth.make_dataset(7, 260, 31)
    folder = tmp_path_factory.mktemp("ds")
    csv = pio.write_dataset(samples, str(folder), binary=True, version="5.1")
    return samples, csv


def test_dataset_statistics_and_batches_match_the_oracle(cuda, disk_dataset):
    from pdivgnn_b200 import io as pio
    samples, csv = disk_dataset
    ds = pio.MeshStressFieldDataset(csv, periodic_graph=True)
    assert len(ds) == 7 and "mesh_filename" in ds.dataframe.columns
    graphs = [O.build_graph(s, True) for s in samples]
    ref = O.dataset_stats(graphs)
    for k, v in ds.stats().items():
        assert abs(float(v) - float(ref[k])) <= 2e-6 * max(1.0, abs(float(ref[k]))), k
    loader = ds.loader(batch_size=3, shuffle=False)
    assert len(loader) == 3
    seen = []
    for j, b in enumerate(loader):
        ids = b.sample_ids
        seen += ids
        ob = O.collate([graphs[i] for i in ids])
        assert torch.equal(b.edge_index.cpu(), ob.edge_index)  # bit-exact index work
        assert torch.equal(b.edge_attr.cpu(), ob.edge_attr)
        assert torch.equal(b.ptr.cpu(), ob.ptr) and torch.equal(b.batch.cpu(), ob.batch)
        assert torch.equal(b.local_stress.cpu(), ob.local_stress)
        assert torch.equal(b.nodes_types.cpu(), ob.nodes_types)
        assert torch.equal(b.op_div_matrix.coalesce().indices().cpu(), ob.op_div_matrix.indices())
    assert seen == list(range(7))
    # sharded + shuffled: every sample exactly once over the ranks, different order per epoch
    l0, l1 = ds.loader(2, shuffle=True, rank=0, world=2, prefetch=False), ds.loader(2, shuffle=True, rank=1, world=2, prefetch=False)
    e0 = sorted(i for b in l0 for i in b.sample_ids) + sorted(i for b in l1 for i in b.sample_ids)
    assert sorted(e0) == list(range(7))


@pytest.mark.parametrize("with_op", [True, False])
def test_resident_dataset_batches_equal_host_batches(cuda, disk_dataset, with_op):
    """batcher.ResidentDataset (whole set in HBM, batches gathered on the device) yields the same MeshBatch, bit for bit,
    as the host collation path, for shuffled ragged batches; both loaders are exercised explicitly."""
    from pdivgnn_b200 import io as pio
    samples, csv = disk_dataset
    ds = pio.MeshStressFieldDataset(csv, periodic_graph=True)
    la = ds.loader(3, shuffle=True, seed=5, with_op_div=with_op, resident=True)
    lb = ds.loader(3, shuffle=True, seed=5, with_op_div=with_op, resident=False)
    assert la.store is not None and lb.store is None and la.store.nbytes() > 0
    n = 0
    for a, b in zip(la, lb):
        assert a.sample_ids == b.sample_ids and a.batch_size == b.batch_size and a.num_nodes == b.num_nodes
        for k in ("pos", "edge_index", "edge_attr", "mean_stress", "local_stress", "nodes_types", "ptr", "batch"):
            assert torch.equal(getattr(a, k), getattr(b, k)), k
        if with_op:
            oa, ob = a.op_div_matrix.coalesce(), b.op_div_matrix.coalesce()
            assert oa.shape == ob.shape and torch.equal(oa.indices(), ob.indices()) and torch.equal(oa.values(), ob.values())
        else:
            assert a.op_div_matrix is None and b.op_div_matrix is None
        n += 1
    assert n == 3


def test_evaluate_and_train_epoch_match_the_oracle(cuda, disk_dataset):
    import pdivgnn_b200
    from pdivgnn_b200 import io as pio
    samples, csv = disk_dataset
    ds = pio.MeshStressFieldDataset(csv, periodic_graph=True)
    sd = O.init_state_dict(seed=69)
    model = H.make_model(ds.stats(), params=sd)
    graphs = [O.build_graph(s, True) for s in samples]
    stats = O.dataset_stats(graphs)
    ev = pdivgnn_b200.evaluate(model, ds.loader(4), monitor_divergence=True)
    ref_tot, ref_nmse = 0.0, 0.0
    for ids in ([0, 1, 2, 3], [4, 5, 6]):
        tot, nmse, div, _ = O.train_loss(sd, O.collate([graphs[i] for i in ids]), stats, 10, True, 1.0, dtype=torch.float64)
        ref_tot += float(tot) / 2
        ref_nmse += float(nmse) / 2
    assert ev["batches"] == 2
    assert abs(ev["total"] - ref_tot) < 2e-5 * abs(ref_tot) and abs(ev["nmse"] - ref_nmse) < 2e-5 * abs(ref_nmse)
    # a few epochs of training through the loader reduce the loss; FusedAdam == the reference's optimizer
    opt = pdivgnn_b200.FusedAdam(model.parameters(), lr=1e-3)
    first = pdivgnn_b200.train_epoch(model, ds.loader(4, shuffle=True), opt, True, 10.0)
    for _ in range(4):
        last = pdivgnn_b200.train_epoch(model, ds.loader(4, shuffle=True), opt, True, 10.0)
    assert first["batches"] == 2 and last["total"] < first["total"]
    assert abs(first["total"] - (first["nmse"] + first["divergence"])) < 1e-5 * abs(first["total"])


def test_predict_and_save_writes_the_reference_layout(cuda, disk_dataset, tmp_path):
    import pdivgnn_b200
    from pdivgnn_b200 import io as pio
    samples, csv = disk_dataset
    ds = pio.MeshStressFieldDataset(csv, periodic_graph=True)
    sd = O.init_state_dict(seed=69)
    model = H.make_model(ds.stats(), params=sd)
    out = pdivgnn_b200.predict_and_save(model, ds.loader(3), str(tmp_path / "res"))
    assert len(out) == 7 and out[3].endswith("fields/hole_plate_mesh_3.npz")
    graphs = [O.build_graph(s, True) for s in samples]
    stats = O.dataset_stats(graphs)
    ref = O.forward(sd, O.collate(graphs[3:6]), stats, 10, scale_output=True, dtype=torch.float64)
    n3 = graphs[3].pos.shape[0]
    z = np.load(out[3])
    org = np.load(ds.dataframe["data_filename"][3])
    assert set(z.files) == set(org.files) and z["stress_field"].shape == (n3, 3)
    assert np.array_equal(z["node_labels"], org["node_labels"])
    linf, l2 = H.rel_err(z["stress_field"], ref[:n3])
    assert linf < 1e-5 and l2 < 1e-5
    with pytest.raises(RuntimeError, match="SHUFFLED"):
        pdivgnn_b200.predict_and_save(model, ds.loader(3, shuffle=True, seed=1), str(tmp_path / "res2"))
