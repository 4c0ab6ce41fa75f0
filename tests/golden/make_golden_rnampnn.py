"""Generate tests/golden/rnampnn_*.npz by running the UNMODIFIED reference ``rnampnn`` package on CPU.

Run once in the build container (needs /root/reference; the GPU box has none):

    python tests/golden/make_golden_rnampnn.py

Same recipe as ``make_golden.py`` (stub modules for the absent third-party packages, seeded
parameters from ``oracle/weights.py`` loaded ``strict=True`` into the reference ``RNAMPNN``).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import REF, install_reference_shims  # noqa: E402
from oracle.weights import RNAMPNNConfig, make_rnampnn_state_dict, synthetic_backbones  # noqa: E402

CFG_FIELDS = list(RNAMPNNConfig.__dataclass_fields__)


def load_reference_model(cfg: RNAMPNNConfig, seed: int):
    install_reference_shims()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from rnampnn.model.rnampnn import RNAMPNN
    model = RNAMPNN(**cfg.as_kwargs())
    res = model.load_state_dict(make_rnampnn_state_dict(cfg, seed), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return model.eval()


def toy():
    """The hand-written tensors of the reference's test.py:12-24 (N=6, one padded row, repeated atoms)."""
    coords = torch.tensor([[
        [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [1.0, 1.0, 0.0], [0.0, 1.0, 1.0], [1.0, 1.0, 0.0], [1.0, 1.0, 0.0], [1.0, 1.0, 0.0]],
        [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 0.0], [1.0, 1.0, 0.0], [2.0, 1.0, 1.732], [1.0, 0.0, 1.0], [1.0, 0.0, 1.0]],
        [[0.0, 1.0, 0.0], [2.0, 1.0, 1.732], [1.0, 0.0, 0.0], [1.0, 0.0, 1.0], [0.0, 0.0, 1.0], [1.0, 0.0, 1.0], [1.0, 0.0, 1.0]],
        [[0.0, 1.0, 1.0], [1.0, 0.0, 0.0], [0.0, 0.0, 0.0], [2.0, 1.0, 1.732], [1.0, 0.0, 1.0], [0.0, 0.0, 1.0], [1.0, 0.0, 1.0]],
        [[0.0, 1.0, 1.0], [2.0, 1.0, 1.732], [0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [1.0, 0.0, 1.0], [0.0, 0.0, 1.0], [1.0, 0.0, 1.0]],
        [[0.0, 0.0, 0.0]] * 7,
    ]])
    mask = torch.tensor([[1., 1., 1., 1., 1., 0.]])
    return coords, mask


def real_structures(names):
    """Seven atoms, NaN -> 0 (the reference's dataset fills NaNs with unseeded random offsets,
    rnampnn/utils/data.py:189-222; zeros keep the fixture reproducible), zero padding, prefix mask."""
    coords = [np.load(os.path.join(REF, "data", "coords", n + ".npy")) for n in names]
    L = max(len(c) for c in coords)
    X = np.zeros((len(names), L, 7, 3), np.float32)
    mask = np.zeros((len(names), L), np.float32)
    for i, c in enumerate(coords):
        X[i, :len(c)] = np.nan_to_num(c, nan=0.0)
        mask[i, :len(c)] = 1
    return torch.from_numpy(X), torch.from_numpy(mask)


@torch.no_grad()
def run_case(name, cfg, seed, coords, mask, keep_layers, edge_stride=1):
    model = load_reference_model(cfg, seed)
    rf = model.res_feature
    edge_index = rf._get_res_graph(coords, mask)
    cap = {}
    h1 = rf.res_edge_embedding_layers.register_forward_hook(lambda m, i, o: cap.__setitem__("edge_raw", i[0].clone()))
    raw, h0, e0, ei2 = rf(coords, mask)
    h1.remove()
    raw, h0, e0 = raw.clone(), h0.clone(), e0.clone()
    assert torch.equal(edge_index, ei2)
    layer_h = []
    hooks = [l.register_forward_hook(lambda m, i, o: layer_h.append(o[0].clone())) for l in model.res_mpnn_layers]
    fused_cap = {}
    h2 = model.post_fusion.register_forward_hook(lambda m, i, o: fused_cap.__setitem__("fused", o.clone()))
    h3 = model.raw_embedding.register_forward_hook(lambda m, i, o: fused_cap.__setitem__("remb", o.clone()))
    e_last = {}
    h4 = model.res_mpnn_layers[-1].register_forward_hook(lambda m, i, o: e_last.__setitem__("e", o[1].clone()))
    logits = model(coords, mask).clone()
    for h in hooks + [h2, h3, h4]:
        h.remove()
    probs = torch.softmax(logits, -1)
    B, N, K = edge_index.shape
    out = dict(
        coords=coords.numpy(), mask=mask.numpy(),
        cfg=np.array([getattr(cfg, f) for f in CFG_FIELDS], np.int64), seed=np.int64(seed),
        edge_stride=np.int64(edge_stride),
        centroids=coords.mean(dim=2).numpy(), edge_index=edge_index.numpy(),
        raw=raw.numpy(), edge_raw=cap["edge_raw"].reshape(B * N * K, -1)[::edge_stride].numpy(),
        h0=h0.numpy(), e0=e0.reshape(B * N * K, -1)[::edge_stride].numpy(),
        layer_ids=np.array(keep_layers, np.int64),
        layers=np.stack([layer_h[i].numpy() for i in keep_layers]),
        e_last=e_last["e"].reshape(B * N * K, -1)[::edge_stride].numpy(),
        fused=fused_cap["fused"].numpy(), raw_embedding=fused_cap["remb"].numpy(),
        logits=logits.numpy(), probs=probs.numpy(), samples=probs.argmax(-1).numpy(),
    )
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: B={B} N={N} K={K} -> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    torch.set_num_threads(4)
    # A: reference toy (test.py), K = L: the "slot == valid_neighbours" quirk, small model
    coords, mask = toy()
    run_case("rnampnn_toy", RNAMPNNConfig(num_res_neighbours=5, num_res_mpnn_layers=2,
                                          num_post_fusion_attn_layers=1), 7, coords, mask, [0, 1])
    # B: ragged batch incl. a structure shorter than K and one of length 1
    coords, _, mask = synthetic_backbones(4, 40, 3001, lengths=[40, 23, 3, 1], atoms=7)
    run_case("rnampnn_ragged", RNAMPNNConfig(num_res_mpnn_layers=2), 11, coords, mask, [0, 1])
    # C: default model (10 layers, 2 attention layers), two structures
    coords, _, mask = synthetic_backbones(2, 64, 3002, lengths=[64, 50], atoms=7)
    run_case("rnampnn_default", RNAMPNNConfig(), 42, coords, mask, [0, 4, 9], edge_stride=3)
    # D: real structures, default model
    coords, mask = real_structures(["2BNY_1_S", "1A9N_1_Q", "1EG0_1_M", "1H2C_1_R"])
    run_case("rnampnn_real", RNAMPNNConfig(), 42, coords, mask, [0, 9], edge_stride=3)
    # E: other shapes: six atoms per feature family (test.py:26-27), K=5, attention in the embedding, B=1
    coords, _, mask = synthetic_backbones(1, 30, 3003, atoms=7)
    run_case("rnampnn_alt", RNAMPNNConfig(num_res_neighbours=5, num_inside_dist_atoms=6, num_cross_dist_atoms=6,
                                          num_embedding_attn_layers=1, num_res_mpnn_layers=1,
                                          num_post_fusion_attn_layers=1, num_readout_layers=1,
                                          num_raw_ffn_layers=2), 13, coords, mask, [0])
    # F: K > N (edge_index padded with -1 columns)
    coords, _, mask = synthetic_backbones(2, 3, 3004, lengths=[3, 2], atoms=7)
    run_case("rnampnn_tiny", RNAMPNNConfig(num_res_neighbours=5, num_res_mpnn_layers=1,
                                           num_post_fusion_attn_layers=1), 17, coords, mask, [0])


if __name__ == "__main__":
    main()
