"""Drop-in ``rdesign.model.feature.RNAFeatures`` (reference: rdesign/model/feature.py:8-238).

kNN graph on P atoms, geometric node/edge features, embeddings and ``Normalize`` --
all on the sm_100a kernels behind the C ABI: a tiled brute-force kNN with a warp-level
top-K (no [B,N,N] tensor), per-node / per-edge feature kernels that recompute their few
distances from coordinates, and the fused embedding + Normalize launch.

Neighbour order: ascending distance, exact ties broken by the lower residue index (the
reference's ``torch.topk`` leaves tie order unspecified).  Only the default feature sets
(node: angle, distance, direction; edge: orientation, distance, direction) are
kernelised -- they are the only ones ``RNAModel`` ever builds.
"""
from __future__ import annotations

import torch
from torch import nn

from rmpnn_b200 import ops
from ..config.glob import FEAT_DIMS
from .functional import Normalize, _kernel_path

_NODE_ORDER = ['angle', 'distance', 'direction']
_EDGE_ORDER = ['orientation', 'distance', 'direction']


class RNAFeatures(nn.Module):
    def __init__(self, edge_features, node_features, node_feat_types=[], edge_feat_types=[], num_rbf=16, top_k=30,
                 augment_eps=0., dropout=0.1):
        super().__init__()
        self.edge_features, self.node_features = edge_features, node_features
        self.top_k, self.augment_eps, self.num_rbf = top_k, augment_eps, num_rbf
        self.dropout = nn.Dropout(dropout)
        self.node_feat_types, self.edge_feat_types = node_feat_types, edge_feat_types
        node_in = sum(FEAT_DIMS['node'][f] for f in node_feat_types)
        edge_in = sum(FEAT_DIMS['edge'][f] for f in edge_feat_types)
        self.node_embedding = nn.Linear(node_in, node_features, bias=True)
        self.edge_embedding = nn.Linear(edge_in, edge_features, bias=True)
        self.norm_nodes = Normalize(node_features)
        self.norm_edges = Normalize(edge_features)

    # -- pieces the reference exposes as methods --------------------------------------
    def _dist(self, X, mask, eps=1E-6):
        """(D_neighbors, E_idx) [B,N,min(top_k,N)] from backbone coordinates X [B,N,3]."""
        _kernel_path(X, mask)
        return ops.knn(X.unsqueeze(2), mask, self.top_k, atom=0)

    def _select(self, raw, order, types, dims):
        """Columns of the full raw feature matrix that the configured feature types use."""
        if list(types) == order:
            return raw
        off, cols = {}, 0
        for name in order:
            off[name] = (cols, cols + dims[name])
            cols += dims[name]
        return torch.cat([raw[:, off[t][0]:off[t][1]] for t in order if t in types], dim=-1)

    def forward(self, X, S, mask):
        _kernel_path(X, mask)
        if self.num_rbf != 16:
            raise ValueError("the B200 feature kernels are built for num_rbf = 16")
        if self.training and self.augment_eps > 0:
            X = X + self.augment_eps * torch.randn_like(X)
        B, N = X.shape[:2]
        g = ops.build_graph(X, mask, self.top_k)
        node_raw, edge_raw = ops.raw_features(X, g)
        node_raw = self._select(node_raw, _NODE_ORDER, self.node_feat_types, FEAT_DIMS['node'])
        edge_raw = self._select(edge_raw, _EDGE_ORDER, self.edge_feat_types, FEAT_DIMS['edge'])
        grad = not _kernel_path(*self.parameters(), training=self.training)
        if grad or self.node_features != ops.HIDDEN or self.edge_features != ops.HIDDEN:
            h_V = self.norm_nodes(self.node_embedding(node_raw))
            h_E = self.norm_edges(self.edge_embedding(edge_raw))
        else:
            h_V = torch.empty(g.Nv, ops.HIDDEN, dtype=torch.float32, device=X.device)
            h_E = torch.empty(g.Ne, ops.HIDDEN, dtype=torch.float32, device=X.device)
            for raw, lin, norm, out, rows in ((node_raw, self.node_embedding, self.norm_nodes, h_V, g.Nv),
                                             (edge_raw, self.edge_embedding, self.norm_edges, h_E, g.Ne)):
                ops.row_mlp([(raw.contiguous(), None)], [(ops.pack_linear(lin.weight, lin.bias), False)],
                            ops.EPI_NORMALIZE, out, rows, gain=norm.gain.detach(), beta=norm.bias.detach())
        src = g.node_src[:g.Nv].long()                       # flat node -> dense b*N+i (index plumbing)
        X_valid = X.reshape(B * N, *X.shape[2:])[src]
        S_valid = S.reshape(B * N)[src]
        batch_id = src // N
        self.last_graph = g
        return X_valid, S_valid, h_V, h_E, g.E_flat, batch_id
