"""pdivgnn_b200 -- B200-native (sm_100a) P-DivGNN message-passing hot path.

Importing the package never touches CUDA; the C-ABI library ``lib/libpdivgnn.so`` is
loaded on first use and the product path raises if it is missing (no CPU fallback).
"""
__version__ = "0.1.0"


def __getattr__(name):  # lazy: keep `import pdivgnn_b200` free of torch/CUDA side effects
    if name in ("EncodeProcessDecode", "Processor", "GraphLayerNorm", "save_model_checkpoint",
                "load_model_checkpoint", "load_optimizer_checkpoint", "print_model"):
        from . import models
        return getattr(models, name)
    if name in ("MeshStressFieldDataset", "read_legacy_vtk", "write_legacy_vtk", "read_sample", "write_sample",
                "write_dataset"):
        from . import io
        return getattr(io, name)
    if name in ("train_epoch", "evaluate", "predict", "predict_and_save"):
        from . import engine
        return getattr(engine, name)
    if name == "FusedAdam":
        from .optim import FusedAdam
        return FusedAdam
    if name == "nmse_div_loss":
        from .loss import nmse_div_loss
        return nmse_div_loss
    raise AttributeError(name)
