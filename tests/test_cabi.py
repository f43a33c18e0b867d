"""CPU checks of the boundary: the C-ABI library loads and exports every symbol
include/pdg.h declares; the module mirrors the reference's state_dict layout."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "pdg.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pdg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from pdivgnn_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    l = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(l, s)]
    assert not missing, missing
    assert set(_lib.declared_symbols()) <= set(syms) | {"pdg_last_error"}


def test_size_queries_without_gpu():
    from pdivgnn_b200 import _lib
    L = _lib.lib()
    assert L.pdg_version() >= 100
    assert L.pdg_plan_bytes(1000, 5600) > 5600 * 4 * 3
    a = L.pdg_forward_ws_bytes(1000, 5600, 10, 0)
    b = L.pdg_forward_ws_bytes(1000, 5600, 10, _lib.FLAG_SAVE)
    assert 0 < a < b
    assert L.pdg_forward_ws_bytes(1000, 5600, 0, 0) == 0  # invalid step count


def test_persistent_grid_is_balanced():
    """pdg_persistent_grid: never more CTAs than SMs, never more rounds than a full grid, every tile covered."""
    from pdivgnn_b200 import _lib
    L = _lib.lib()
    assert L.pdg_persistent_grid(1516, 148) == 138 and L.pdg_persistent_grid(260, 148) == 130  # configs[1]
    assert L.pdg_persistent_grid(0, 148) == 0 and L.pdg_persistent_grid(5, 148) == 5 and L.pdg_persistent_grid(148, 148) == 148
    for sms in (1, 7, 132, 148):
        for n in list(range(1, 700)) + [1516, 8300, 46000]:
            g = L.pdg_persistent_grid(n, sms)
            rounds_full = -(-n // sms)
            assert 1 <= g <= min(n, sms)
            assert -(-n // g) == rounds_full          # same number of rounds as `sms` CTAs
            assert g == -(-n // rounds_full)          # ... and no smaller grid achieves it
            assert g * rounds_full >= n


def test_state_dict_layout_and_init():
    import pdivgnn_b200
    from oracle import pdg_oracle as O
    torch.manual_seed(69)
    m = pdivgnn_b200.EncodeProcessDecode(1, 10, 128, 6, 3)
    sd = m.state_dict()
    assert list(sd.keys()) == O.STATE_KEYS
    assert sum(v.numel() for v in sd.values()) == 167299
    ref = O.init_state_dict(seed=69)
    assert all(torch.equal(sd[k], ref[k]) for k in O.STATE_KEYS)
    assert sd["processor.edge_net.0.weight"].shape == (128, 384)
    assert sd["processor.node_net.0.weight"].shape == (128, 256)
    assert sd["node_decoder.2.weight"].shape == (3, 128)
    for a in ("latent_size", "message_passing_steps", "input_nodes_features_size", "input_edges_features_size",
              "output_nodes_features_size", "mean_pos", "std_local_stress"):
        assert hasattr(m, a)


def test_no_cpu_fallback():
    import pdivgnn_b200
    from types import SimpleNamespace
    m = pdivgnn_b200.EncodeProcessDecode(1, 2, 128, 6, 3)
    g = SimpleNamespace(mean_stress=torch.ones(4, 3), pos=torch.zeros(4, 2), nodes_types=torch.zeros(4, 1, dtype=torch.long),
                        edge_index=torch.tensor([[0, 1], [1, 0]]), edge_attr=torch.ones(2))
    with pytest.raises(RuntimeError, match="CUDA"):
        m(g)
    with pytest.raises(NotImplementedError):
        pdivgnn_b200.EncodeProcessDecode(1, 2, 64, 6, 3)


def test_checkpoint_roundtrip(tmp_path):
    import pdivgnn_b200
    from pdivgnn_b200 import models
    torch.manual_seed(1)
    kw = dict(mean_pos=torch.tensor(50.), std_pos=torch.tensor(29.), mean_mean_stress=torch.tensor(1.),
              std_mean_stress=torch.tensor(2.), mean_local_stress=torch.tensor(3.), std_local_stress=torch.tensor(4.),
              mean_edge_weight=torch.tensor(5.), std_edge_weight=torch.tensor(6.))
    m = pdivgnn_b200.EncodeProcessDecode(1, 10, 128, 6, 3, **kw)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    f = str(tmp_path / "ck.pth")
    models.save_model_checkpoint(m, opt, 7, f)
    ck = torch.load(f)
    assert set(ck.keys()) == {"model_state_dict", "optimizer_state_dict", "epoch", "mean_pos", "mean_mean_stress",
                              "std_mean_stress", "mean_local_stress", "std_pos", "std_local_stress",
                              "mean_edge_weight", "std_edge_weight"}
    m2 = pdivgnn_b200.EncodeProcessDecode(1, 10, 128, 6, 3)
    assert models.load_model_checkpoint(m2, f) == 7
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    assert float(m2.std_edge_weight) == 6.0
    n = m2._norm_struct()
    assert (n.mean_pos, n.std_pos, n.std_local_stress) == (50.0, 29.0, 4.0)


def test_per_channel_statistics_are_rejected_not_truncated():
    """the reference broadcasts per-channel statistics; the C ABI carries scalars -- refuse, do not take [0]."""
    import pdivgnn_b200
    m = pdivgnn_b200.EncodeProcessDecode(1, 10, 128, 6, 3, mean_pos=torch.tensor([1.0, 2.0]), std_pos=torch.tensor(1.0),
                                         mean_mean_stress=torch.tensor(0.0), std_mean_stress=torch.tensor(1.0),
                                         mean_local_stress=torch.tensor(0.0), std_local_stress=torch.tensor(1.0),
                                         mean_edge_weight=torch.tensor(0.0), std_edge_weight=torch.tensor(1.0))
    with pytest.raises(NotImplementedError, match="mean_pos"):
        m._norm_struct()
    m.mean_pos = (3.0,)  # a 1-tuple is a scalar
    assert m._norm_struct().mean_pos == 3.0


def test_ctypes_signatures_match_the_header():
    """Every binding in pdivgnn_b200._lib takes as many arguments as include/pdg.h declares, pointer arguments are bound as
    pointers and 64-bit integers as 64-bit (a short argument list or an int where the header says int64_t would corrupt
    the call silently)."""
    import ctypes as C
    from pdivgnn_b200 import _lib
    txt = open(os.path.join(ROOT, "include", "pdg.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    decls = {m.group(1): m.group(2) for m in re.finditer(r"\b(pdg_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", txt)}
    checked = 0
    for name, (_res, args) in _lib._SIGS.items():
        assert name in decls, f"{name} is bound but not declared in include/pdg.h"
        params = [p.strip() for p in decls[name].split(",")]
        if params == ["void"] or params == [""]:
            params = []
        assert len(params) == len(args), (name, len(params), len(args))
        for p, a in zip(params, args):
            is_ptr = "*" in p
            bound_ptr = a is C.c_void_p or a is C.c_char_p or (isinstance(a, type) and issubclass(a, C._Pointer))
            assert is_ptr == bound_ptr, (name, p, a)
            if not is_ptr:
                base = p.replace("const", "").split()[0]
                if base in ("int64_t", "size_t"):
                    assert C.sizeof(a) == 8, (name, p, a)
                elif base in ("int", "int32_t"):
                    assert C.sizeof(a) == 4, (name, p, a)
                elif base == "double":
                    assert a is C.c_double, (name, p, a)
                elif base == "float":
                    assert a is C.c_float, (name, p, a)
        checked += 1
    assert checked >= 30
