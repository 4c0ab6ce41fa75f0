"""The reference-facing module API (rnampnn.model.*) of the B200 build.

CPU part: constructor / state_dict compatibility with the reference's names, loud failure without CUDA.
GPU part (-m gpu): each drop-in module against the oracle and the golden vectors of the unmodified reference."""
import inspect

import numpy as np
import pytest
import torch

from oracle import rnampnn_oracle as O
from oracle.weights import RNAMPNNConfig, make_rnampnn_state_dict, synthetic_backbones
from conftest import assert_close_fp32, to_fp64


def _model(cfg, seed, device=None):
    from rnampnn.model.rnampnn import RNAMPNN
    m = RNAMPNN(**cfg.as_kwargs())
    res = m.load_state_dict(make_rnampnn_state_dict(cfg, seed), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return m.to(device) if device else m


# ------------------------------------------------------------------ CPU
def test_state_dict_names_match_reference_checkpoints():
    m = _model(RNAMPNNConfig(), 42)
    assert sum(p.numel() for p in m.parameters()) == 3_536_900        # reference RNAMPNN(), probed
    sd = m.state_dict()
    for k, shape in (("res_feature.raw_project.weight", (128, 28)),
                     ("res_feature.res_embedding.ffn_layers.9.weight", (128, 512)),
                     ("res_feature.graph_norm.scale", (1, 1, 128)),
                     ("res_feature.res_edge_embedding_layers.0.weight", (128, 90)),
                     ("res_mpnn_layers.9.message_layers.3.bias", (128,)),
                     ("res_mpnn_layers.0.edge_layers.0.weight", (128, 384)),
                     ("post_fusion.bi_attention_layers.1.in_proj_weight", (384, 128)),
                     ("post_fusion.bi_attention_layers.0.out_proj.bias", (128,)),
                     ("post_fusion.graph_norm_layers.1.shift", (1, 1, 128)),
                     ("raw_embedding.raw_ffn.0.weight", (512, 28)),
                     ("raw_embedding.graph_norm.scale", (1, 1, 128)),
                     ("readout.readout_layers.0.weight", (512, 256)),
                     ("readout.readout_layers.3.weight", (4, 512))):
        assert tuple(sd[k].shape) == shape, k
    assert m.hparams.num_res_neighbours == 3 and m.hparams.padding_len == 4500 and m.hparams.dropout == 0.4


def test_module_surface_matches_reference_signatures():
    from rnampnn.model.feature import ResFeature
    from rnampnn.model.functional import GraphNormalization, RawFFN, Readout, RNABert
    from rnampnn.model.mpnn import ResMPNN
    from rnampnn.model.rnampnn import RNAMPNN
    sig = lambda c: list(inspect.signature(c.__init__).parameters)[1:]
    assert sig(ResFeature) == ["num_neighbours", "num_inside_dist_atoms", "num_inside_angle_atoms",
                               "num_inside_dihedral_atoms", "num_cross_dist_atoms", "num_cross_angle_atoms",
                               "num_cross_dihedral_atoms", "res_embedding_dim", "padding_len", "num_attn_layers",
                               "num_heads", "ffn_dim", "num_ffn_layers", "res_edge_embedding_dim", "num_edge_layers",
                               "dropout"]
    assert sig(ResMPNN) == ["res_embedding_dim", "res_edge_embedding_dim", "depth_res_mpnn", "num_edge_layers", "dropout"]
    assert sig(GraphNormalization) == ["embedding_dim"]
    assert sig(Readout) == ["embedding_dim", "readout_hidden_dim", "num_layers", "dropout"]
    assert sig(RNABert) == ["padding_len", "res_embedding_dim", "num_attn_layers", "num_heads", "ffn_dim",
                            "num_ffn_layers", "dropout"]
    assert sig(RawFFN) == ["raw_dim", "num_raw_ffn_dim", "num_raw_ffn_layers", "raw_embedding_dim", "dropout"]
    assert len(sig(RNAMPNN)) == 35
    assert list(inspect.signature(RNAMPNN.forward).parameters)[1:] == ["coords", "mask", "is_predict"]
    assert list(inspect.signature(ResMPNN.forward).parameters)[1:] == ["res_embedding", "res_edge_embedding",
                                                                       "edge_index", "mask"]


def test_no_cpu_path():
    m = _model(RNAMPNNConfig(num_res_mpnn_layers=1), 1).eval()
    coords, _, mask = synthetic_backbones(1, 8, 1, atoms=7)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(coords, mask)


def test_collate_layout():
    from rnampnn.utils.data import featurize, separate
    items = [{"id": "a", "sequence": torch.eye(4)[[0, 1, 2]], "coordinates": torch.ones(3, 7, 3)},
             {"id": "b", "sequence": torch.eye(4)[[3, 3, 1, 0, 2]], "coordinates": 2 * torch.ones(5, 7, 3)}]
    seq, coords, mask, ids = featurize(items)
    assert seq.shape == (2, 5, 4) and coords.shape == (2, 5, 7, 3) and ids == ["a", "b"]
    assert mask.tolist() == [[1, 1, 1, 0, 0], [1, 1, 1, 1, 1]] and float(coords[0, 3:].abs().sum()) == 0
    sep = separate(torch.arange(8.0), mask.sum(-1))
    assert sep.tolist() == [[0, 1, 2, 0, 0], [3, 4, 5, 6, 7]]


# ------------------------------------------------------------------ GPU
gpu = pytest.mark.gpu
SMALL = ("rnampnn_toy", "rnampnn_tiny")


@gpu
def test_model_forward_matches_reference(rgolden):
    """RNAMPNN.forward (eval) from coordinates against the unmodified reference's logits, when the graphs agree."""
    coords, mask, cfg = rgolden["coords"], rgolden["mask"], rgolden["config"]
    m = _model(cfg, int(rgolden["seed"]), "cuda").eval()
    with torch.no_grad():
        logits = m(coords.cuda(), mask.cuda())
        emb = m.embedding(coords.cuda(), mask.cuda())
    assert logits.shape == (*mask.shape, 4) and emb.shape == (*mask.shape, 256)
    ref = O.forward(rgolden["params"], coords, mask, cfg)
    ref64 = O.forward(to_fp64(rgolden["params"]), coords.double(), mask.double(), cfg, edge_index=ref["edge_index"])
    tol = 1e-3 if rgolden["name"] in SMALL else 1e-4
    assert (logits.cpu() - ref["logits"]).abs().max() < tol
    assert_close_fp32(emb, ref["embedding"], ref64["embedding"], tol, "embedding")
    if torch.equal(ref["edge_index"], torch.from_numpy(rgolden["edge_index"])):
        assert np.abs(logits.cpu().numpy() - rgolden["logits"]).max() < tol


@gpu
def test_model_bf16_mode_within_budget():
    """RNAMPNN with precision_mode = "bf16" (RMPNN_PRECISION=bf16): tensor-core Linear stacks, 2e-2 logit budget."""
    cfg = RNAMPNNConfig()
    coords, _, mask = synthetic_backbones(3, 90, 91, lengths=[90, 61, 7], atoms=7)
    m = _model(cfg, 42, "cuda").eval()
    m.precision_mode = "bf16"
    with torch.no_grad():
        logits = m(coords.cuda(), mask.cuda())
    ref = O.forward(make_rnampnn_state_dict(cfg, 42), coords, mask, cfg)
    assert (logits.cpu() - ref["logits"]).abs().max() < 2e-2


@gpu
def test_res_feature_module(rgolden):
    from rnampnn.model.feature import ResFeature
    coords, mask, cfg = rgolden["coords"], rgolden["mask"], rgolden["config"]
    m = _model(cfg, int(rgolden["seed"]), "cuda").eval()
    assert isinstance(m.res_feature, ResFeature)
    with torch.no_grad():
        raw, h0, e0, ei = m.res_feature(coords.cuda(), mask.cuda())
    want = O.res_feature(rgolden["params"], coords, mask, cfg)
    assert torch.equal(ei.cpu(), want[3]) and ei.dtype == torch.int64
    want64 = O.res_feature(to_fp64(rgolden["params"]), coords.double(), mask.double(), cfg, edge_index=want[3])
    atol = 5e-4 if rgolden["name"] in SMALL else 5e-5
    np.testing.assert_allclose(raw.cpu().numpy(), want[0].numpy(), atol=2e-5, rtol=1e-6)
    assert_close_fp32(h0, want[1], want64[1], atol, "res_embedding")
    assert_close_fp32(e0, want[2], want64[2], atol, "res_edge_embedding")
    assert torch.equal(m.res_feature._get_res_graph(coords.cuda(), mask.cuda()).cpu(), want[3])


@gpu
def test_res_mpnn_module_standalone_and_in_place_edges():
    """ResMPNN.forward on caller-made tensors (reference's graph with a -1 sentinel row), e updated in place."""
    from rnampnn.model.mpnn import ResMPNN
    cfg = RNAMPNNConfig(num_res_mpnn_layers=1)
    params = make_rnampnn_state_dict(cfg, 5)
    coords, _, mask = synthetic_backbones(2, 30, 77, lengths=[30, 2], atoms=7)
    _, h, e, ei = O.res_feature(params, coords, mask, cfg)
    want_h, want_e = O.res_mpnn_layer(params, 0, h, e, ei, mask)
    layer = ResMPNN(128, 128, 2, 2, 0.4)
    layer.load_state_dict({k[len("res_mpnn_layers.0."):]: v for k, v in params.items() if k.startswith("res_mpnn_layers.0.")})
    layer = layer.cuda().eval()
    e_dev = e.clone().cuda()
    with torch.no_grad():
        msg = layer.message(h.cuda(), e_dev, ei.cuda(), mask.cuda())
        agg = layer.aggregation(mask.cuda(), h.cuda(), msg, ei.cuda())
        got_h, got_e = layer(h.cuda(), e_dev, ei.cuda(), mask.cuda())
    assert msg.shape == e.shape and float(msg[ei.cuda() == -1].abs().max()) == 0.0
    assert agg.shape == h.shape
    np.testing.assert_allclose(got_h.cpu().numpy(), want_h.numpy(), atol=5e-5, rtol=1e-5)
    np.testing.assert_allclose(got_e.cpu().numpy(), want_e.numpy(), atol=5e-5, rtol=1e-5)
    assert got_e.data_ptr() == e_dev.data_ptr()                         # in place like the reference (mpnn.py:263)


@gpu
def test_functional_modules_standalone():
    from rnampnn.model.functional import GraphNormalization, RawFFN, Readout, RNABert
    g = torch.Generator().manual_seed(3)
    B, N = 3, 50
    mask = torch.ones(B, N); mask[1, 20:] = 0; mask[2, 1:] = 0
    x = torch.randn(B, N, 128, generator=g)
    cfg = RNAMPNNConfig()
    params = make_rnampnn_state_dict(cfg, 9)
    sub = lambda pre: {k[len(pre):]: v for k, v in params.items() if k.startswith(pre)}
    with torch.no_grad():
        gn = GraphNormalization(128); gn.load_state_dict(sub("raw_embedding.graph_norm.")); gn = gn.cuda().eval()
        want = O.graph_norm(x, mask, params["raw_embedding.graph_norm.scale"], params["raw_embedding.graph_norm.shift"])
        np.testing.assert_allclose(gn(x.cuda(), mask.cuda()).cpu().numpy(), want.numpy(), atol=2e-5, rtol=1e-5)

        bert = RNABert(4500, 128, 2, 8, 512, 3, 0.4); bert.load_state_dict(sub("post_fusion.")); bert = bert.cuda().eval()
        want = O.rna_bert(params, "post_fusion.", x * mask[..., None], mask, 8, 4500)
        np.testing.assert_allclose(bert(x.cuda() * mask.cuda()[..., None], mask.cuda()).cpu().numpy(), want.numpy(),
                                   atol=1e-4, rtol=1e-5)

        raw = torch.rand(B, N, 28, generator=g) * 5
        ffn = RawFFN(28, 512, 3, 128, 0.4); ffn.load_state_dict(sub("raw_embedding.")); ffn = ffn.cuda().eval()
        np.testing.assert_allclose(ffn(raw.cuda(), mask.cuda()).cpu().numpy(), O.raw_ffn(params, raw, mask).numpy(),
                                   atol=1e-4, rtol=1e-5)

        emb = torch.randn(B, N, 256, generator=g)
        ro = Readout(256, 512, 2, 0.4); ro.load_state_dict(sub("readout.")); ro = ro.cuda().eval()
        got = ro(emb.cuda(), mask.cuda()).cpu()
        np.testing.assert_allclose(got.numpy(), O.readout(params, emb, mask).numpy(), atol=5e-5, rtol=1e-5)
        assert float(got[mask == 0].abs().max()) == 0.0


@gpu
def test_training_mode_composes_the_same_math_and_differentiates():
    """train() + autograd: the differentiable torch-CUDA composition agrees with the kernel path (dropout 0)."""
    cfg = RNAMPNNConfig(num_res_mpnn_layers=2)
    from rnampnn.model.rnampnn import RNAMPNN
    m = RNAMPNN(**cfg.as_kwargs(), dropout=0.0)
    m.load_state_dict(make_rnampnn_state_dict(cfg, 21), strict=True)
    m = m.cuda()
    coords, _, mask = synthetic_backbones(2, 40, 88, lengths=[40, 17], atoms=7)
    m.eval()
    with torch.no_grad():
        ref = m(coords.cuda(), mask.cuda())
    m.train()
    out = m(coords.cuda(), mask.cuda())
    assert out.requires_grad
    assert (out - ref).abs().max() < 2e-4
    out.square().sum().backward()
    for name, p in m.named_parameters():
        if name.startswith("res_mpnn_layers.1.edge_layers."):
            continue                                    # the last layer's edge update never reaches the logits
        assert p.grad is not None and torch.isfinite(p.grad).all(), name


@gpu
def test_non_default_widths_use_the_composition():
    """test.py:79 builds ResMPNN with 32-wide embeddings: widths the kernels are not built for still run (torch-CUDA)."""
    from rnampnn.model.mpnn import ResMPNN
    layer = ResMPNN(res_embedding_dim=32, res_edge_embedding_dim=32, depth_res_mpnn=2, num_edge_layers=2, dropout=0.1)
    layer = layer.cuda().eval()
    h = torch.randn(1, 6, 32).cuda(); e = torch.randn(1, 6, 3, 32).cuda()
    ei = torch.tensor([[[1, 2, -1], [0, 2, 3], [1, 3, 0], [2, 4, 1], [3, 2, -1], [-1, -1, -1]]]).cuda()
    mask = torch.tensor([[1., 1., 1., 1., 1., 0.]]).cuda()
    with torch.no_grad():
        h2, e2 = layer(h, e, ei, mask)
    assert h2.shape == h.shape and e2.shape == e.shape and float(h2[0, 5].abs().max()) == 0.0
