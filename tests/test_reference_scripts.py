"""Script-level drop-in: with ``rna-mpnn_b200/`` first on ``sys.path`` every name the reference's
entry scripts import resolves, and their call chains run.

* ``test.py:1-6`` imports ``AtomMPNN``, ``AtomFeature``, ``to_atom_format``, ``RNADataModule`` next to the hot-path
  modules; its three live tests (``test_atom_feature`` :34, ``test_atom_mpnn`` :87, ``test_recovery`` :218) are
  restated here on the reference's own toy coordinates (the reference file cannot travel to the GPU box).
* ``main.py:17-31``: ``seeding() -> RNAModel -> load_state_dict -> predict(model, batch_size)`` over a competition
  directory; ``train.py:62-80``: ``seeding(); RNAModel(); RNADataModule.from_defaults(...); get_trainer(...)``.
"""
import inspect
import os

import numpy as np
import pytest
import torch


def toy_residues():
    """The six residues of the reference's toy tensors (test.py:12-24): integer-grid atoms with repeats (zero bond
    vectors) and an all-zero padded residue."""
    P = {"o": [0.0, 0.0, 0.0], "x": [1.0, 0.0, 0.0], "xy": [1.0, 1.0, 0.0], "yz": [0.0, 1.0, 1.0], "y": [0.0, 1.0, 0.0],
         "t": [2.0, 1.0, 1.732], "xz": [1.0, 0.0, 1.0], "z": [0.0, 0.0, 1.0]}
    rows = [["o", "x", "xy", "yz", "xy", "xy", "xy"], ["o", "x", "o", "xy", "t", "xz", "xz"],
            ["y", "t", "x", "xz", "z", "xz", "xz"], ["yz", "x", "o", "t", "xz", "z", "xz"],
            ["yz", "t", "o", "x", "xz", "z", "xz"], ["o"] * 7]
    return torch.tensor([[P[a] for a in r] for r in rows])


# ------------------------------------------------------------------------------------------ imports (test.py:1-6)
def test_every_name_the_entry_scripts_import_resolves():
    from rnampnn.model.mpnn import ResMPNN, AtomMPNN                      # noqa: F401  test.py:1
    from rnampnn.model.feature import ResFeature, AtomFeature, to_atom_format   # noqa: F401  test.py:2
    from rnampnn.model.functional import Readout                         # noqa: F401  test.py:3 (the two Bert* names
    from rnampnn.model.rnampnn import RNAMPNN                            # noqa: F401   no longer exist upstream either)
    from rnampnn.utils.data import RNADataModule                         # noqa: F401  test.py:6
    from rnampnn.model.pool import AtomPool, ResPool                     # noqa: F401
    from rnampnn.utils.train import get_trainer as rt, LossMonitor, NameModel, XGBTrainer   # noqa: F401
    from rnampnn.utils.predict import predict as rp                      # noqa: F401
    from rnampnn.utils.seed import seeding as rs                         # noqa: F401
    from rdesign.model.rdesign import RNAModel                           # noqa: F401  main.py:17
    from rdesign.config.glob import BEST_CKPT, DATA_PATH, OUTPUT_PATH    # noqa: F401  main.py:18
    from rdesign.config.seeds import seeding                             # noqa: F401  main.py:19
    from rdesign.utils.predict import predict                            # noqa: F401  main.py:20
    from rdesign.utils.data import RNADataModule as RD                   # noqa: F401  train.py:64
    from rdesign.utils.train import get_trainer                          # noqa: F401  train.py:65
    assert list(inspect.signature(get_trainer).parameters) == ["name", "version", "max_epochs", "val_check_interval",
                                                               "train_xgb"]
    assert inspect.signature(get_trainer).parameters["max_epochs"].default == 30
    assert list(inspect.signature(rt).parameters) == ["name", "version", "max_epochs", "val_check_interval",
                                                      "progress_bar"]
    assert inspect.signature(rt).parameters["max_epochs"].default == 60
    assert hasattr(LossMonitor, "on_validation_epoch_end") and hasattr(LossMonitor, "on_test_epoch_end")
    assert hasattr(XGBTrainer, "on_fit_end") and hasattr(XGBTrainer, "_generate_embedding")


# ------------------------------------------------------------------------------------------ test.py:34, :87 (CPU ok)
def test_atom_feature_like_reference_test():
    from rnampnn.model.feature import AtomFeature, to_atom_format
    res = toy_residues()
    coords = torch.stack([res[0], res[5]])[None]                          # one real residue + one padded
    mask = torch.tensor([[1, 0]])
    atom_coords, atom_mask = to_atom_format(coords, mask)
    assert atom_coords.shape == (1, 14, 3) and atom_mask.tolist() == [[1] * 7 + [0] * 7]
    emb, dists, index = AtomFeature(num_atom_neighbours=7)(atom_coords, atom_mask)
    assert emb.shape == (1, 14, 128) and dists.shape == (1, 14, 7) and index.shape == (1, 14, 7)
    assert (index[0, 7:] == -1).all() and (dists[0, 7:] == 1e6).all() and (emb[0, 7:] == 0).all()
    # atom 2 coincides with atoms 4, 5, 6: its three nearest neighbours are exactly those, at sqrt(1e-6)
    assert sorted(index[0, 2, :3].tolist()) == [4, 5, 6] and torch.allclose(dists[0, 2, :3], torch.tensor(1e-3))
    assert not (index[0, :7] == torch.arange(7)[:, None]).any()           # never its own neighbour


def test_atom_mpnn_like_reference_test():
    from rnampnn.model.feature import AtomFeature, to_atom_format
    from rnampnn.model.mpnn import AtomMPNN
    res = toy_residues()
    coords = torch.stack([res[0], res[5]])[None].repeat(2, 1, 1, 1)
    mask = torch.tensor([[1, 0], [1, 0]])
    atom_coords, atom_mask = to_atom_format(coords, mask)
    feat, mpnn = AtomFeature(num_atom_neighbours=5, atom_embedding_dim=32), AtomMPNN(atom_embedding_dim=32, depth_atom_mpnn=2)
    emb, dists, index = feat(atom_coords, atom_mask)
    msg = mpnn.message(emb, dists, index, atom_mask)
    out = mpnn(emb, dists, index, atom_mask)
    assert msg.shape == (2, 14, 5, 32) and out.shape == (2, 14, 32)
    assert (msg[:, 7:] == 0).all() and (out[:, 7:] == 0).all() and torch.isfinite(out).all()
    out.sum().backward()
    assert all(p.grad is not None for p in mpnn.parameters())


def test_atom_pool_shapes():
    from rnampnn.model.pool import AtomPool
    pool = AtomPool(raw_dim=28, atom_pool_hidden_dim=16, num_layers=3, dropout=0.1).eval()
    out = pool(torch.randn(2, 3 * 7, 32), torch.ones(2, 3 * 7), torch.randn(2, 3, 28))
    assert out.shape == (2, 3, 32)


# ------------------------------------------------------------------------------------------ data module (test.py:148)
def _write_dataset(root, lengths, with_nan=True):
    rng = np.random.default_rng(5)
    os.makedirs(os.path.join(root, "coords")); os.makedirs(os.path.join(root, "seqs"))
    for i, L in enumerate(lengths):
        c = rng.normal(size=(L, 7, 3)) * 10
        if with_nan and i % 2 == 0:
            c[0, 3] = np.nan; c[L - 1, 6] = np.nan
        np.save(os.path.join(root, "coords", f"rna{i}.npy"), c)
        with open(os.path.join(root, "seqs", f"rna{i}.fasta"), "w") as f:
            f.write(f">rna{i} test\n" + "".join("AUCG"[j % 4] for j in range(L)) + "\n")


def test_rnampnn_data_module(tmp_path):
    from rnampnn.utils.data import RNADataModule, RNADataset
    _write_dataset(str(tmp_path), [5, 9, 12, 7, 30, 3, 8, 11, 6, 10])
    dm = RNADataModule(data_path=str(tmp_path), split_ratio=0.8, batch_size=4, noise_augmentation=3)
    dm.num_workers = 0
    dm.setup("fit")
    assert len(dm.train_set) + len(dm.val_set) == 13
    assert not set(dm.train_set.get_ids()) & set(dm.val_set.get_ids())     # augmented copies stay with their source
    sequences, coords, mask, ids = next(iter(dm.train_dataloader()))
    assert sequences.shape[:2] == coords.shape[:2] == mask.shape and coords.shape[2:] == (7, 3) and len(ids) == 4
    assert torch.isfinite(coords).all() and (sequences.sum(-1) == mask).all()
    ds = RNADataset.from_path(str(tmp_path), is_predict=True)
    item = ds.getitem_by_key("rna0")[0]
    assert item["sequence"][:, 0].all() and item["coordinates"].shape == (5, 7, 3)
    # the parked atoms sit 1.5 A / 4 A from their anchors (data.py:189-222)
    raw = np.load(os.path.join(str(tmp_path), "coords", "rna0.npy"))
    assert abs(float((item["coordinates"][0, 3] - torch.tensor(raw[0, 0], dtype=torch.float32)).norm()) - 1.5) < 1e-4
    assert abs(float((item["coordinates"][4, 6] - item["coordinates"][4, 5]).norm()) - 4.0) < 1e-4
    ds.slice_augmentation(2, min_len=4)
    assert len(ds) == 12 and ds[-1]["sequence"].shape[0] == 4


def test_get_trainer_needs_lightning_only_at_call_time():
    from rdesign.utils.train import get_trainer
    try:
        import pytorch_lightning  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="pytorch_lightning"):
            get_trainer("RDesign-X", 3, max_epochs=1, train_xgb=True)


# ------------------------------------------------------------------------------------------ test.py:218 (GPU)
@pytest.mark.gpu
def test_recovery_like_reference_test():
    """``RNAMPNN()(coords, mask)`` on the two toy structures with masks 11111 0 / 1111 00 (the second keeps real
    coordinates under a zero mask), default 10-layer model, integer mask -- logits finite, masked rows zero, and the
    same logits as the CPU oracle of the reference path with the same seeded weights."""
    from oracle import rnampnn_oracle as RO
    from oracle.weights import RNAMPNNConfig
    from rnampnn.model.rnampnn import RNAMPNN
    res = toy_residues()
    coords = torch.stack([res, res])
    mask = torch.tensor([[1, 1, 1, 1, 1, 0], [1, 1, 1, 1, 0, 0]])
    torch.manual_seed(0)
    model = RNAMPNN().cuda().eval()
    logits = model(coords.cuda(), mask.cuda())
    assert logits.shape == (2, 6, 4) and torch.isfinite(logits).all()
    assert (logits[0, 5] == 0).all() and (logits[1, 4:] == 0).all()
    params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref = RO.forward(params, coords, mask.float(), RNAMPNNConfig())
    probs = torch.softmax(logits, -1)[mask.bool().cuda()]
    assert probs.shape == (9, 4)
    # exact ties at 1e6 between "self" and padded residues (L_b <= K) are broken by the lower index on both sides
    assert (logits.cpu() - ref["logits"]).abs().max().item() < 1e-4


@pytest.mark.gpu
def test_main_py_call_chain(tmp_path, monkeypatch):
    """main.py:25-31 against the drop-in: seeding -> RNAModel -> state_dict (stands in for the absent checkpoint)
    -> predict(model, batch_size=4) over a competition directory -> submit.csv with one sequence per structure."""
    import pandas as pd
    from rdesign.config.seeds import seeding
    from rdesign.model.rdesign import RNAModel
    from rdesign.utils.predict import predict
    from oracle.weights import RDesignConfig, make_state_dict
    data, out = tmp_path / "saisdata", tmp_path / "saisresult"
    lengths = [40, 33, 57, 12, 80]
    _write_dataset(str(data), lengths)
    out.mkdir()
    monkeypatch.chdir(tmp_path)
    seeding()
    model = RNAModel().cuda()
    model.load_state_dict(make_state_dict(RDesignConfig(), 42), strict=True)
    try:
        predict(model, batch_size=4, data_path=str(data) + "/", output_path=str(out) + "/", num_workers=0)
    except torch.OutOfMemoryError:                                         # main.py:30-31
        predict(model, batch_size=2, data_path=str(data) + "/", output_path=str(out) + "/", num_workers=0)
    df = pd.read_csv(out / "submit.csv")
    assert sorted(df["pdb_id"]) == sorted(f"rna{i}" for i in range(len(lengths)))
    by_id = dict(zip(df["pdb_id"], df["seq"]))
    assert all(len(by_id[f"rna{i}"]) == L and set(by_id[f"rna{i}"]) <= set("AUCG") for i, L in enumerate(lengths))


@pytest.mark.gpu
def test_train_py_constructor_chain(tmp_path, monkeypatch):
    """train.py:67-79 up to the trainer: seeding(); RNAModel(); RNADataModule.from_defaults(batch_size, split_ratio);
    one training_step + configure_optimizers on a batch of that module (Lightning itself is not in this image)."""
    from rdesign.config.seeds import seeding
    from rdesign.model.rdesign import RNAModel
    from rdesign.utils.data import RNADataModule
    _write_dataset(str(tmp_path / "data"), [30, 22, 41, 18, 25, 36], with_nan=False)
    monkeypatch.chdir(tmp_path)
    seeding()
    model = RNAModel(num_mpnn_layers=2).cuda()
    dm = RNADataModule.from_defaults(batch_size=3, split_ratio=[0.5, 0.5, 0.0])
    dm.num_workers = 0
    dm.setup("fit")
    batch = next(iter(dm.train_dataloader()))
    batch = [t.cuda() if torch.is_tensor(t) else t for t in batch]
    opt = model.configure_optimizers()
    opt = opt[0][0] if isinstance(opt, tuple) else (opt["optimizer"] if isinstance(opt, dict) else opt)
    model.train()
    loss = model.training_step(batch)
    loss.backward()
    opt.step()
    assert torch.isfinite(loss)
