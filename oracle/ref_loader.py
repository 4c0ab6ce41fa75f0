"""TEST INFRASTRUCTURE -- imports the UNMODIFIED reference (``/root/reference`` in the build container, the staged
copy ``oracle/_ref`` on the GPU box) behind stub modules for the third-party packages this image lacks
(pytorch_lightning, xgboost, Bio, matplotlib, seaborn).  Used by the golden-vector generators, the CPU reference arm
of ``bench.py`` and ``scripts/bench_reference_gpu.py``; never by the product path.

The reference's top-level package names (``rdesign``, ``rnampnn``) are the same as the drop-in's, so the loader
imports the reference under those names with its own root first on ``sys.path`` and then restores ``sys.modules`` /
``sys.path``: callers get the reference classes without disturbing whichever package the process already uses."""
from __future__ import annotations

import os
import sys
import types

import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    """Directory holding the reference's ``rdesign`` package, or None."""
    for root in (os.environ.get("RNAMPNN_REFERENCE", "/root/reference"), os.path.join(HERE, "_ref")):
        if os.path.isfile(os.path.join(root, "rdesign", "model", "rdesign.py")):
            return root
    return None


def install_shims():
    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    try:
        import pytorch_lightning  # noqa: F401
    except ImportError:
        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                import inspect
                loc = inspect.currentframe().f_back.f_locals
                self.hparams = types.SimpleNamespace(**{k: v for k, v in loc.items() if k not in ("self", "__class__")})

            def log(self, *a, **k):
                pass

            @property
            def device(self):
                return next(self.parameters()).device

            def on_load_checkpoint(self, checkpoint):
                pass

        pl = stub("pytorch_lightning", LightningModule=LightningModule,
                  LightningDataModule=type("LightningDataModule", (), {}),
                  Callback=type("Callback", (), {}), Trainer=type("Trainer", (), {}))
        pl.callbacks = stub("pytorch_lightning.callbacks", ModelCheckpoint=type("ModelCheckpoint", (), {}),
                            Callback=pl.Callback)
        pl.loggers = stub("pytorch_lightning.loggers", TensorBoardLogger=type("TensorBoardLogger", (), {}))
        pl.utilities = stub("pytorch_lightning.utilities")
        pl.utilities.types = stub("pytorch_lightning.utilities.types", STEP_OUTPUT=object)
    try:
        import xgboost  # noqa: F401
    except ImportError:
        stub("xgboost", XGBClassifier=type("XGBClassifier", (), {"__init__": lambda self, *a, **k: None}))
    try:
        import Bio  # noqa: F401
    except ImportError:
        bio = stub("Bio")
        bio.SeqIO = stub("Bio.SeqIO")
        bio.PDB = stub("Bio.PDB", PDBParser=object, is_aa=lambda *a, **k: False)
    for name in ("matplotlib", "seaborn"):
        try:
            __import__(name)
        except ImportError:
            stub(name)
    if "matplotlib.pyplot" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except ImportError:
            sys.modules["matplotlib"].pyplot = stub("matplotlib.pyplot")


def load_reference_classes():
    """{'RNAModel': <reference class>, 'RNAMPNN': <reference class>, 'root': path} or None if no reference is found."""
    root = reference_root()
    if root is None:
        return None
    install_shims()
    saved = {k: v for k, v in sys.modules.items() if k == "rdesign" or k.startswith("rdesign.")
             or k == "rnampnn" or k.startswith("rnampnn.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        from rdesign.model.rdesign import RNAModel
        from rnampnn.model.rnampnn import RNAMPNN
        out = {"RNAModel": RNAModel, "RNAMPNN": RNAMPNN, "root": root}
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k == "rdesign" or k.startswith("rdesign.")
                  or k == "rnampnn" or k.startswith("rnampnn.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return out
