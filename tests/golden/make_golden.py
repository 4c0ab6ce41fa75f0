"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run once in the build container (needs /root/reference; the GPU box has none):

    python tests/golden/make_golden.py

The reference's ``rdesign`` package is imported from /root/reference through stub
modules for the packages this image lacks (pytorch_lightning, xgboost, Bio,
matplotlib, seaborn -- SURVEY.md appendix B).  Seeded parameters from
``oracle/weights.py`` are loaded ``strict=True`` into the reference ``RNAModel``;
inputs, the reference's intermediate tensors (captured with forward hooks) and
its outputs are stored.  The committed .npz files are what pins the oracle.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("RNAMPNN_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle.weights import RDesignConfig, make_state_dict, synthetic_backbones  # noqa: E402


def install_reference_shims():
    """Stub modules for absent third-party packages so the reference imports."""
    import torch.nn as nn

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
                frame = inspect.currentframe().f_back
                loc = {k: v for k, v in frame.f_locals.items() if k not in ("self", "__class__")}
                self.hparams = types.SimpleNamespace(**loc)

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
        class XGBClassifier:
            def __init__(self, *a, **k):
                pass
        stub("xgboost", XGBClassifier=XGBClassifier)
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


def load_reference_model(cfg: RDesignConfig, seed: int):
    install_reference_shims()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from rdesign.model.rdesign import RNAModel
    model = RNAModel(**cfg.as_kwargs())
    missing = model.load_state_dict(make_state_dict(cfg, seed), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.eval()


def degenerate_toy():
    """Two tiny structures on an integer grid with repeated atoms: exact distance
    ties, zero-length bond vectors (normalize 0/0 paths), one and two padded rows."""
    a = torch.tensor([
        [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 1], [1, 1, 0], [1, 1, 0]],
        [[0, 0, 0], [1, 0, 0], [0, 0, 0], [1, 1, 0], [2, 1, 2], [1, 0, 1]],
        [[0, 1, 0], [2, 1, 2], [1, 0, 0], [1, 0, 1], [0, 0, 1], [1, 0, 1]],
        [[0, 1, 1], [1, 0, 0], [0, 0, 0], [2, 1, 2], [1, 0, 1], [0, 0, 1]],
        [[3, 1, 1], [2, 1, 2], [0, 0, 0], [1, 0, 0], [1, 0, 1], [0, 0, 1]],
        [[0, 0, 0]] * 6,
    ], dtype=torch.float32)
    b = a.clone().flip(0).roll(1, 0) * 1.5
    b[4:] = 0
    X = torch.stack([a, b])
    mask = torch.tensor([[1, 1, 1, 1, 1, 0], [1, 1, 1, 1, 0, 0]], dtype=torch.float32)
    S = torch.tensor([[0, 1, 2, 3, 0, 0], [3, 2, 1, 0, 0, 0]])
    return X, S, mask


def real_structures(names):
    """Reference collate semantics (rdesign/utils/data.py:84-119): first six atoms,
    NaN -> 0 before padding, residues kept, zero padding, prefix mask."""
    coords = [np.load(os.path.join(REF, "data", "coords", n + ".npy"))[:, :6, :] for n in names]
    seqs = []
    for n in names:
        with open(os.path.join(REF, "data", "seqs", n + ".fasta")) as f:
            seqs.append("".join(l.strip() for l in f if not l.startswith(">")))
    L = max(len(c) for c in coords)
    X = np.zeros((len(names), L, 6, 3), np.float32)
    S = np.zeros((len(names), L), np.int64)
    mask = np.zeros((len(names), L), np.float32)
    for i, (c, s) in enumerate(zip(coords, seqs)):
        X[i, :len(c)] = np.nan_to_num(c, nan=0.0)
        S[i, :len(c)] = ["AUCG".index(ch) for ch in s]
        mask[i, :len(c)] = 1
    return torch.from_numpy(X), torch.from_numpy(S), torch.from_numpy(mask)


@torch.no_grad()
def run_case(name, cfg, seed, X, S, mask, keep_layers, edge_stride=1):
    model = load_reference_model(cfg, seed)
    cap = {}
    feats = model.features
    h1 = feats.node_embedding.register_forward_hook(lambda m, i, o: cap.__setitem__("node_raw", i[0].clone()))
    h2 = feats.edge_embedding.register_forward_hook(lambda m, i, o: cap.__setitem__("edge_raw", i[0].clone()))
    layer_out = []
    hooks = [l.register_forward_hook(lambda m, i, o: layer_out.append(o.clone())) for l in model.mpnn_layers]
    D_nb, E_dense = feats._dist(X[:, :, 0, :], mask)
    Xv, Sv, h_V0, h_E, E_flat, batch_id = feats(X, S, mask)
    h_V, S_out = model(X, S, mask)
    layer_out = layer_out[-len(model.mpnn_layers):]
    logits = model.readout(h_V)
    probs = torch.softmax(logits, -1)
    for h in [h1, h2] + hooks:
        h.remove()
    out = dict(
        X=X.numpy(), S=S.numpy(), mask=mask.numpy(),
        cfg=np.array([cfg.hidden_dim, cfg.k_neighbors, cfg.num_message_layers, cfg.num_dense_layers,
                      cfg.dim_dense_layers, cfg.num_mpnn_layers, cfg.readout_hidden_dim,
                      cfg.num_readout_layers], np.int64),
        seed=np.int64(seed), edge_stride=np.int64(edge_stride),
        D_neighbors=D_nb.numpy(), E_dense=E_dense.numpy(),
        node_raw=cap["node_raw"].numpy(), edge_raw=cap["edge_raw"][::edge_stride].numpy(),
        h_V0=h_V0.numpy(), h_E=h_E[::edge_stride].numpy(), E_idx=E_flat.numpy(),
        batch_id=batch_id.numpy(), S_valid=S_out.numpy(), X_valid=Xv.numpy(),
        h_V=h_V.numpy(), logits=logits.numpy(), probs=probs.numpy(),
        samples=probs.argmax(-1).numpy(),
        layer_ids=np.array(keep_layers, np.int64),
        layers=np.stack([layer_out[i].numpy() for i in keep_layers]),
    )
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: Nv={h_V.shape[0]} Ne={E_flat.shape[1]} -> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    torch.set_num_threads(4)
    # A: degenerate toy, K > some lengths, tiny model
    X, S, mask = degenerate_toy()
    run_case("toy_degenerate", RDesignConfig(k_neighbors=3, num_mpnn_layers=2), 7, X, S, mask, [0, 1])
    # B: ragged batch incl. a structure shorter than K
    X, S, mask = synthetic_backbones(3, 40, 2001, lengths=[40, 23, 5])
    run_case("ragged_k8", RDesignConfig(k_neighbors=8, num_mpnn_layers=2), 11, X, S, mask, [0, 1])
    # C: default 9-layer model (K=25), two structures
    X, S, mask = synthetic_backbones(2, 64, 2002, lengths=[64, 50])
    run_case("default_k25", RDesignConfig(), 42, X, S, mask, [0, 4, 8], edge_stride=5)
    # D: real structures with NaN->0 atoms, default model
    X, S, mask = real_structures(["2BNY_1_S", "1A9N_1_Q", "1EG0_1_M", "1H2C_1_R"])
    run_case("real_k25", RDesignConfig(), 42, X, S, mask, [0, 8], edge_stride=3)
    # E: B=1 (the reference's .squeeze() quirk), K=48, other layer shape options
    X, S, mask = synthetic_backbones(1, 96, 2003)
    run_case("single_k48", RDesignConfig(k_neighbors=48, num_mpnn_layers=1, num_message_layers=2,
                                         num_dense_layers=1, dim_dense_layers=512,
                                         num_readout_layers=2), 13, X, S, mask, [0], edge_stride=9)


if __name__ == "__main__":
    main()
