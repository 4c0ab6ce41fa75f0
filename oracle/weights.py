"""Deterministic parameter sets under the reference's ``state_dict`` names.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

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
