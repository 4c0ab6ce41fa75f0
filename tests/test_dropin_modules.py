"""The reference-facing module API (rdesign.model.*) of the B200 build.

CPU part: constructor / state_dict compatibility with the reference's names, loud
failure without CUDA.  GPU part (-m gpu): each drop-in module against the oracle."""
import numpy as np
import pytest
import torch

from oracle import rdesign_oracle as O
from oracle.weights import RDesignConfig, make_state_dict, synthetic_backbones


def _model(cfg, seed, device=None):
    from rdesign.model.rdesign import RNAModel
    m = RNAModel(**cfg.as_kwargs())
    res = m.load_state_dict(make_state_dict(cfg, seed), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return m.to(device) if device else m


# ------------------------------------------------------------------ CPU
def test_state_dict_names_match_reference_checkpoints():
    cfg = RDesignConfig()
    m = _model(cfg, 42)
    assert sum(p.numel() for p in m.parameters()) == 2_551_812
    keys = set(m.state_dict())
    for k in ("features.node_embedding.weight", "features.norm_edges.gain", "mpnn_layers.8.message_layers.6.bias",
              "mpnn_layers.0.dense.9.weight", "mpnn_layers.3.norm2.weight", "readout.readout_layers.0.weight"):
        assert k in keys
    assert m.state_dict()["features.edge_embedding.weight"].shape == (128, 115)
    assert m.hparams.k_neighbors == 25 and m.hparams.lr == 0.002


def test_module_surface_matches_reference_signatures():
    import inspect
    from rdesign.model.feature import RNAFeatures
    from rdesign.model.functional import Normalize, Readout, scatter_sum
    from rdesign.model.mpnn import MPNNLayer
    assert list(inspect.signature(RNAFeatures.__init__).parameters)[1:] == [
        "edge_features", "node_features", "node_feat_types", "edge_feat_types", "num_rbf", "top_k", "augment_eps",
        "dropout"]
    assert list(inspect.signature(MPNNLayer.__init__).parameters)[1:] == [
        "num_hidden", "num_in", "num_message_layers", "num_dense_layers", "dim_dense_layers", "dropout", "scale"]
    assert list(inspect.signature(Readout.__init__).parameters)[1:] == [
        "embedding_dim", "readout_hidden_dim", "num_layers", "dropout"]
    assert list(inspect.signature(Normalize.__init__).parameters)[1:] == ["features", "epsilon"]
    out = scatter_sum(torch.ones(4, 2), torch.tensor([0, 0, 2, 2]), dim=0)
    assert out.tolist() == [[2, 2], [0, 0], [2, 2]]


def test_no_cpu_path():
    m = _model(RDesignConfig(num_mpnn_layers=1), 1).eval()
    X, S, mask = synthetic_backbones(1, 8, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(X, S, mask)


def test_collate_layout():
    from rdesign.utils.data import featurize, gen_mask, separate, concat
    rng = np.random.default_rng(0)
    c1 = rng.normal(size=(5, 7, 3)); c1[0, 0] = np.nan
    c2 = rng.normal(size=(3, 7, 3))
    atoms = ["P", "O5'", "C5'", "C4'", "C3'", "O3'"]
    batch = [{"name": "a", "seq": "AUCGA", "coords": {a: c1[:, i] for i, a in enumerate(atoms)}},
             {"name": "b", "seq": "GGC", "coords": {a: c2[:, i] for i, a in enumerate(atoms)}}]
    X, S, mask, lengths, names = featurize(batch)
    assert X.shape == (2, 5, 6, 3) and X.dtype == torch.float32 and S.dtype == torch.int64
    assert torch.equal(X[0, 0, 0], torch.zeros(3))                     # NaN atom -> origin, residue kept
    assert mask.tolist() == [[1, 1, 1, 1, 1], [1, 1, 1, 0, 0]] and lengths.tolist() == [5, 3]
    assert S[1].tolist() == [3, 3, 2, 0, 0] and names == ["a", "b"]
    flat = torch.arange(8.)
    assert torch.equal(concat(separate(flat, lengths), lengths), flat)
    assert torch.equal(gen_mask(lengths), mask)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_rnamodel_forward_matches_reference_golden(golden):
    cfg = golden["config"]
    m = _model(cfg, int(golden["seed"]), "cuda").eval()
    X, S, mask = golden["X"].cuda(), golden["S"].cuda(), golden["mask"].cuda()
    with torch.no_grad():
        h_V, S_v = m(X, S, mask)
        logits = m.readout(h_V)
    assert np.array_equal(S_v.cpu().numpy(), golden["S_valid"])
    if golden["X"].shape[1] == 6:
        return            # degenerate integer-lattice toy: dihedral signs are rounding-order dependent
    ref = O.forward(golden["params"], golden["X"], golden["S"], golden["mask"], cfg.k_neighbors, cfg.scale)
    np.testing.assert_allclose(logits.cpu().numpy(), ref["logits"].numpy(), atol=1e-4, rtol=0)
    assert torch.equal(logits.argmax(-1).cpu(), ref["samples"])
    # eval mode uses the kernels even with autograd on (outputs carry no graph)
    h2, _ = m(X, S, mask)
    assert not h2.requires_grad and torch.equal(h2, h_V)


@pytest.mark.gpu
def test_features_module_standalone(golden):
    from rdesign.model.feature import RNAFeatures
    cfg = golden["config"]
    f = RNAFeatures(128, 128, node_feat_types=['angle', 'distance', 'direction'],
                    edge_feat_types=['orientation', 'distance', 'direction'], top_k=cfg.k_neighbors)
    f.load_state_dict({k[len("features."):]: v for k, v in golden["params"].items() if k.startswith("features.")})
    f = f.cuda().eval()
    Xv, Sv, h_V, h_E, E_idx, batch_id = f(golden["X"].cuda(), golden["S"].cuda(), golden["mask"].cuda())
    ref = O.features(golden["params"], golden["X"], golden["S"], golden["mask"], cfg.k_neighbors)
    for a, b in zip((Xv, Sv, E_idx, batch_id), (ref[0], ref[1], ref[4], ref[5])):
        assert torch.equal(a.cpu(), b)
    if golden["X"].shape[1] != 6:   # see tests/test_cuda_parity.py::FRAGILE_CASE
        np.testing.assert_allclose(h_V.cpu().numpy(), ref[2].numpy(), atol=5e-5, rtol=0)
        np.testing.assert_allclose(h_E.cpu().numpy(), ref[3].numpy(), atol=2e-4, rtol=0)
    D, E = f._dist(golden["X"][:, :, 0, :].cuda(), golden["mask"].cuda())
    D_o, E_o = O.knn(golden["X"][:, :, 0, :], golden["mask"], cfg.k_neighbors)
    assert torch.equal(E.cpu(), E_o) and torch.equal(D.cpu(), D_o)


@pytest.mark.gpu
def test_mpnn_layer_standalone_with_concatenated_input():
    from rdesign.model.mpnn import MPNNLayer
    cfg = RDesignConfig(k_neighbors=6, num_mpnn_layers=1)
    params = make_state_dict(cfg, 9)
    X, S, mask = synthetic_backbones(2, 20, 3, lengths=[20, 9])
    _, _, h_V, h_E, E_idx, _ = O.features(params, X, S, mask, 6)
    ref = O.mpnn_layer(params, 0, h_V, h_E, E_idx)
    layer = MPNNLayer(128, 256, 3, 3, 256)
    layer.load_state_dict({k[len("mpnn_layers.0."):]: v for k, v in params.items() if k.startswith("mpnn_layers.0.")})
    layer = layer.cuda().eval()
    h_EV = torch.cat([h_E, h_V[E_idx[0]], h_V[E_idx[1]]], -1).cuda()
    out = layer(h_V.cuda(), h_EV, E_idx.cuda())
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=5e-5, rtol=0)
    # shuffled edge order gives the same sums (atomics make the segmented sum order-free)
    perm = torch.randperm(E_idx.shape[1])
    out2 = layer(h_V.cuda(), h_EV[perm.cuda()], E_idx[:, perm].cuda())
    np.testing.assert_allclose(out2.cpu().numpy(), ref.numpy(), atol=5e-5, rtol=0)


@pytest.mark.gpu
def test_normalize_and_readout_standalone():
    from rdesign.model.functional import Normalize, Readout
    g = torch.Generator().manual_seed(0)
    x = torch.randn(77, 128, generator=g) * 3 + 1
    n = Normalize(128)
    with torch.no_grad():
        n.gain.copy_(torch.randn(128, generator=g)); n.bias.copy_(torch.randn(128, generator=g))
    ref = O.normalize_rows(x, n.gain.detach(), n.bias.detach())
    out = n.cuda().eval()(x.cuda())
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=2e-5, rtol=0)
    for layers in (0, 1, 3):
        r = Readout(128, 256, layers)
        ref = O.readout({"readout." + k: v.detach() for k, v in r.state_dict().items()}, x)
        out = r.cuda().eval()(x.cuda())
        np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=2e-5, rtol=0)


@pytest.mark.gpu
def test_training_mode_composes_autograd_path():
    cfg = RDesignConfig(k_neighbors=5, num_mpnn_layers=2)
    from rdesign.model.rdesign import RNAModel
    m = RNAModel(dropout=0.0, **{k: v for k, v in cfg.as_kwargs().items()})
    m.load_state_dict(make_state_dict(cfg, 2))
    m = m.cuda()
    X, S, mask = synthetic_backbones(2, 16, 6, lengths=[16, 7])
    m.eval()
    with torch.no_grad():
        h_eval, _ = m(X.cuda(), S.cuda(), mask.cuda())
    m.train()
    h_train, S_v = m(X.cuda(), S.cuda(), mask.cuda())
    assert h_train.requires_grad
    np.testing.assert_allclose(h_train.detach().cpu().numpy(), h_eval.cpu().numpy(), atol=2e-4, rtol=0)
    loss = m.training_step((X, S, mask, np.array([16, 7], np.int32), ["a", "b"]))
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())


@pytest.mark.gpu
def test_predict_writes_reference_csv(tmp_path):
    cfg = RDesignConfig(k_neighbors=7, num_mpnn_layers=2)
    m = _model(cfg, 4, "cuda")
    X, S, mask = synthetic_backbones(3, 25, 12, lengths=[25, 11, 4])
    lengths = np.array([25, 11, 4], np.int32)
    m.predict((X, S, mask, lengths, ["s0", "s1", "s2"]), 0, str(tmp_path), "out.csv")
    rows = (tmp_path / "out.csv").read_text().strip().splitlines()
    assert rows[0] == "pdb_id,seq"
    ref = O.forward(make_state_dict(cfg, 4), X, S, mask, 7)["samples"].tolist()
    want = "".join("AUCG"[i] for i in ref)
    got = "".join(r.split(",")[1] for r in rows[1:])
    assert got == want and [len(r.split(",")[1]) for r in rows[1:]] == [25, 11, 4]
