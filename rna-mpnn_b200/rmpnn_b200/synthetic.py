"""Workload definition shared by the benches, the tests and the oracle: deterministic parameter sets under the
reference's ``state_dict`` names, the model configurations, and the synthetic backbones of SURVEY.md section 8d.
Pure data generation (seeded torch generators) -- no algorithm of the path lives here, which is why it sits in the
product package and not under ``oracle/`` (``oracle/weights.py`` re-exports it for the tests).

The reference ships no trained checkpoint (``/root/reference/.MISSING_LARGE_BLOBS``),
so parity is checked with seeded weights loaded ``strict=True`` into both the
reference ``RNAModel`` (when generating ``tests/golden``) and the B200 drop-in.
Names/shapes follow ``rdesign/model/rdesign.py:50-65`` (module tree),
``rdesign/model/feature.py:24-27`` (embeddings + Normalize),
``rdesign/model/mpnn.py:11-29`` (LayerNorms, message / dense Sequentials whose
indices skip the GELU/Dropout slots) and ``rdesign/model/functional.py:111-121``
(readout).  Values are drawn per tensor from its own ``torch.Generator`` so they
do not depend on module construction order.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, asdict

import torch

NODE_IN = 12 + 80 + 9      # rdesign/config/glob.py:8-13  angle | distance | direction
EDGE_IN = 4 + 96 + 15      # rdesign/config/glob.py:14-18 orientation | distance | direction


@dataclass(frozen=True)
class RDesignConfig:
    """Constructor arguments of ``RNAModel`` that shape the forward path
    (``rdesign/model/rdesign.py:20-38``)."""
    hidden_dim: int = 128
    k_neighbors: int = 25
    num_message_layers: int = 3
    num_dense_layers: int = 3
    dim_dense_layers: int = 256
    num_mpnn_layers: int = 9
    readout_hidden_dim: int = 256
    num_readout_layers: int = 0
    scale: float = 30.0        # rdesign/model/mpnn.py:6

    def as_kwargs(self) -> dict:
        d = asdict(self)
        d.pop("scale")
        return d


def _linear(sd, name, out_f, in_f, gen_seed):
    g = torch.Generator().manual_seed(gen_seed)
    bound = 1.0 / (in_f ** 0.5)
    sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
    sd[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * bound


def _affine(sd, wname, bname, n, gen_seed):
    g = torch.Generator().manual_seed(gen_seed)
    sd[wname] = 1.0 + 0.1 * torch.randn(n, generator=g)
    sd[bname] = 0.1 * torch.randn(n, generator=g)


def make_state_dict(cfg: RDesignConfig = RDesignConfig(), seed: int = 42) -> "OrderedDict[str, torch.Tensor]":
    """fp32 parameters for ``RNAModel(**cfg.as_kwargs())`` under its state_dict keys."""
    H = cfg.hidden_dim
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    ctr = [seed * 1_000_003]

    def nxt():
        ctr[0] += 1
        return ctr[0]

    _linear(sd, "features.node_embedding", H, NODE_IN, nxt())
    _linear(sd, "features.edge_embedding", H, EDGE_IN, nxt())
    _affine(sd, "features.norm_nodes.gain", "features.norm_nodes.bias", H, nxt())
    _affine(sd, "features.norm_edges.gain", "features.norm_edges.bias", H, nxt())
    for l in range(cfg.num_mpnn_layers):
        p = f"mpnn_layers.{l}"
        _affine(sd, f"{p}.norm1.weight", f"{p}.norm1.bias", H, nxt())
        _affine(sd, f"{p}.norm2.weight", f"{p}.norm2.bias", H, nxt())
        in_f = 3 * H
        for m in range(cfg.num_message_layers):
            _linear(sd, f"{p}.message_layers.{3 * m}", H, in_f, nxt())
            in_f = H
        # rdesign/model/mpnn.py:22-28: `input_dim` carries over from the message loop
        for d in range(cfg.num_dense_layers):
            _linear(sd, f"{p}.dense.{3 * d}", cfg.dim_dense_layers, in_f, nxt())
            in_f = cfg.dim_dense_layers
        _linear(sd, f"{p}.dense.{3 * cfg.num_dense_layers}", H, in_f, nxt())
    in_f = H
    for r in range(max(cfg.num_readout_layers - 1, 0)):
        _linear(sd, f"readout.readout_layers.{3 * r}", cfg.readout_hidden_dim, in_f, nxt())
        in_f = cfg.readout_hidden_dim
    _linear(sd, f"readout.readout_layers.{3 * max(cfg.num_readout_layers - 1, 0)}", 4, in_f, nxt())
    return sd


def synthetic_backbones(B: int, L: int, seed: int, lengths=None, atoms: int = 6):
    """Random-walk RNA backbones: P = cumsum(3.5*randn),
    atoms = P + 1.5*randn; prefix masks from ``lengths``; zero padding like
    ``featurize`` (``rdesign/utils/data.py:110-118``)."""
    g = torch.Generator().manual_seed(seed)
    P = torch.cumsum(3.5 * torch.randn(B, L, 1, 3, generator=g), dim=1)
    X = P + 1.5 * torch.randn(B, L, atoms, 3, generator=g)
    S = torch.randint(0, 4, (B, L), generator=g)
    mask = torch.ones(B, L)
    if lengths is not None:
        for b, n in enumerate(lengths):
            mask[b, n:] = 0.0
        X = X * mask[:, :, None, None]
        S = S * mask.long()
    return X.float().contiguous(), S.long(), mask.float()


# =========================================================================== rnampnn (secondary path)
@dataclass(frozen=True)
class RNAMPNNConfig:
    """Constructor arguments of ``RNAMPNN`` that shape the forward path
    (``rnampnn/model/rnampnn.py:19-54``; defaults are the reference's)."""
    num_res_neighbours: int = 3
    num_inside_dist_atoms: int = 7
    num_inside_angle_atoms: int = 6
    num_inside_dihedral_atoms: int = 6
    num_cross_dist_atoms: int = 7
    num_cross_angle_atoms: int = 6
    num_cross_dihedral_atoms: int = 6
    res_embedding_dim: int = 128
    num_embedding_attn_layers: int = 0
    num_embedding_heads: int = 8
    embedding_ffn_dim: int = 512
    num_embedding_ffn_layers: int = 3
    res_edge_embedding_dim: int = 128
    depth_res_edge_feature: int = 2
    num_res_mpnn_layers: int = 10
    depth_res_mpnn: int = 2
    num_mpnn_edge_layers: int = 2
    padding_len: int = 4500
    num_post_fusion_attn_layers: int = 2
    num_post_fusion_heads: int = 8
    post_fusion_ffn_dim: int = 512
    num_post_fusion_ffn_layers: int = 3
    num_raw_ffn_dim: int = 512
    num_raw_ffn_layers: int = 3
    raw_embedding_dim: int = 128
    readout_hidden_dim: int = 512
    num_readout_layers: int = 2

    def as_kwargs(self) -> dict:
        return asdict(self)

    @property
    def raw_dim(self) -> int:          # rnampnn/model/feature.py:181
        n = self.num_inside_dist_atoms
        return n * (n - 1) // 2 + self.num_inside_angle_atoms - 2 + self.num_inside_dihedral_atoms - 3

    @property
    def raw_edge_dim(self) -> int:     # rnampnn/model/feature.py:182
        return (self.num_cross_dist_atoms ** 2 + (self.num_cross_angle_atoms - 1) ** 2
                + (self.num_cross_dihedral_atoms - 2) ** 2)


def _graph_norm(sd, prefix, n, gen_seed):
    g = torch.Generator().manual_seed(gen_seed)
    sd[prefix + ".scale"] = (1.0 + 0.1 * torch.randn(n, generator=g)).reshape(1, 1, n)
    sd[prefix + ".shift"] = (0.1 * torch.randn(n, generator=g)).reshape(1, 1, n)


def _bert(sd, prefix, H, n_attn, ffn_dim, n_ffn, nxt):
    """``RNABert`` parameters (rnampnn/model/functional.py:105-127)."""
    for l in range(n_attn):
        g = torch.Generator().manual_seed(nxt())
        bound = 1.0 / (H ** 0.5)
        sd[f"{prefix}.bi_attention_layers.{l}.in_proj_weight"] = (torch.rand(3 * H, H, generator=g) * 2 - 1) * bound
        sd[f"{prefix}.bi_attention_layers.{l}.in_proj_bias"] = (torch.rand(3 * H, generator=g) * 2 - 1) * 0.1
        _linear(sd, f"{prefix}.bi_attention_layers.{l}.out_proj", H, H, nxt())
    for l in range(n_attn):
        _graph_norm(sd, f"{prefix}.graph_norm_layers.{l}", H, nxt())
    in_f = H
    for l in range(n_ffn):
        _linear(sd, f"{prefix}.ffn_layers.{3 * l}", ffn_dim, in_f, nxt())
        in_f = ffn_dim
    _linear(sd, f"{prefix}.ffn_layers.{3 * n_ffn}", H, ffn_dim, nxt())


def make_rnampnn_state_dict(cfg: RNAMPNNConfig = RNAMPNNConfig(), seed: int = 42):
    """fp32 parameters for ``RNAMPNN(**cfg.as_kwargs())`` under its state_dict keys
    (module tree: rnampnn/model/rnampnn.py:94-134, feature.py:184-203, mpnn.py:131-152,
    functional.py:62-74, :105-127, :179-189)."""
    H, He = cfg.res_embedding_dim, cfg.res_edge_embedding_dim
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    ctr = [seed * 1_000_003 + 500_000]

    def nxt():
        ctr[0] += 1
        return ctr[0]

    _linear(sd, "res_feature.raw_project", H, cfg.raw_dim, nxt())
    _bert(sd, "res_feature.res_embedding", H, cfg.num_embedding_attn_layers, cfg.embedding_ffn_dim,
          cfg.num_embedding_ffn_layers, nxt)
    _graph_norm(sd, "res_feature.graph_norm", H, nxt())
    in_f = cfg.raw_edge_dim
    for l in range(cfg.depth_res_edge_feature):
        _linear(sd, f"res_feature.res_edge_embedding_layers.{3 * l}", He, in_f, nxt())
        in_f = He
    for i in range(cfg.num_res_mpnn_layers):
        p = f"res_mpnn_layers.{i}"
        _graph_norm(sd, f"{p}.graph_norm", H, nxt())
        in_f = 2 * H + He
        for l in range(cfg.depth_res_mpnn):
            _linear(sd, f"{p}.message_layers.{3 * l}", H, in_f, nxt())
            in_f = H
        in_f = 2 * H + He
        for l in range(cfg.num_mpnn_edge_layers):
            _linear(sd, f"{p}.edge_layers.{3 * l}", He, in_f, nxt())
            in_f = He
    _bert(sd, "post_fusion", H, cfg.num_post_fusion_attn_layers, cfg.post_fusion_ffn_dim,
          cfg.num_post_fusion_ffn_layers, nxt)
    in_f = cfg.raw_dim
    for l in range(cfg.num_raw_ffn_layers):
        _linear(sd, f"raw_embedding.raw_ffn.{3 * l}", cfg.num_raw_ffn_dim, in_f, nxt())
        in_f = cfg.num_raw_ffn_dim
    _linear(sd, f"raw_embedding.raw_ffn.{3 * cfg.num_raw_ffn_layers}", cfg.raw_embedding_dim, in_f, nxt())
    _graph_norm(sd, "raw_embedding.graph_norm", cfg.raw_embedding_dim, nxt())
    in_f = cfg.raw_embedding_dim + H
    for l in range(cfg.num_readout_layers - 1):
        _linear(sd, f"readout.readout_layers.{3 * l}", cfg.readout_hidden_dim, in_f, nxt())
        in_f = cfg.readout_hidden_dim
    _linear(sd, f"readout.readout_layers.{3 * (cfg.num_readout_layers - 1)}", 4, in_f, nxt())
    return sd
