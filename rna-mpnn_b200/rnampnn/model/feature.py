"""Drop-in ``rnampnn.model.feature.ResFeature`` (reference: rnampnn/model/feature.py:133-592).

Residue graph on atom centroids (self excluded, -1 sentinels), 28 intra-residue and 90 cross-residue raw
features, and their embeddings -- on the sm_100a kernels behind the C ABI: the tiled brute-force kNN with a
warp-level top-K (no [B,N,N] tensor), per-residue / per-edge feature kernels that recompute their few
distances from coordinates, and the fused row-MLP for the Linear stacks.

Neighbour order: ascending centroid distance, exact ties broken by the lower residue index (the reference's
``torch.topk`` leaves tie order unspecified).

``AtomFeature`` / ``to_atom_format`` (reference :10-130) are never reached from ``RNAMPNN.forward``; they are kept as
plain-PyTorch passthroughs (any device) so that ``test.py::test_atom_*`` and user code importing them keep working.
"""
from __future__ import annotations

import torch
from torch import nn

from rmpnn_b200 import ops
from rmpnn_b200.engine_rnampnn import RNAMPNNEngine
from ..config.glob import DEFAULT_HIDDEN_DIM, LEPS, NUM_MAIN_SEQ_ATOMS, SEPS
from .functional import GraphNormalization, RNABert, _dropout_active, _flat, _kernel_path, _packed, _widths_ok


def to_atom_format(coords: torch.Tensor, mask: torch.Tensor):
    """Residue layout -> atom layout: coords [B,N,A,3] -> [B,N*A,3], mask [B,N] -> [B,N*7] (every residue's flag
    repeated for its seven atoms)."""
    B = coords.shape[0]
    return coords.reshape(B, -1, coords.shape[3]), mask.repeat_interleave(NUM_MAIN_SEQ_ATOMS, dim=1)


class AtomFeature(nn.Module):
    """Atom-level graph and type embedding (passthrough, plain PyTorch; O(M^2) in the number of atoms like the
    reference).  Neighbours = the ``num_atom_neighbours`` nearest atoms by sqrt(|d|^2 + 1e-6); the atom itself and
    pairs with a padded atom sit at 1e6; the self slot becomes -1, and a padded atom's whole row is -1 / 1e6."""

    def __init__(self, num_atom_neighbours: int, atom_embedding_dim: int = DEFAULT_HIDDEN_DIM):
        super().__init__()
        self.num_atom_neighbours = num_atom_neighbours
        self.atom_embedding_dim = atom_embedding_dim
        self.embedding = nn.Embedding(num_embeddings=NUM_MAIN_SEQ_ATOMS, embedding_dim=atom_embedding_dim)
        self.graph_norm = GraphNormalization(embedding_dim=atom_embedding_dim)

    def _get_atom_graph(self, atom_coords: torch.Tensor, atom_mask: torch.Tensor):
        B, M, _ = atom_coords.shape
        K = self.num_atom_neighbours
        m = atom_mask.to(atom_coords.dtype)
        delta = atom_coords[:, None, :, :] - atom_coords[:, :, None, :]
        dist = torch.sqrt(delta.square().sum(dim=-1) + SEPS)
        dist = dist + LEPS * torch.eye(M, device=dist.device, dtype=dist.dtype)
        both = m[:, None, :] * m[:, :, None]
        dist = dist * both + (1.0 - both) * LEPS
        near_d, near_i = torch.topk(dist, min(K, M), dim=-1, largest=False)
        if K > M:                                              # fewer atoms than requested neighbours
            near_d = torch.cat([near_d, near_d.new_full((B, M, K - M), LEPS)], dim=-1)
            near_i = torch.cat([near_i, near_i.new_full((B, M, K - M), -1)], dim=-1)
        own = torch.arange(M, device=near_i.device)[None, :, None]
        near_i = torch.where(near_i == own, torch.full_like(near_i, -1), near_i)
        padded = (atom_mask == 0)[:, :, None]
        return near_d.masked_fill(padded, LEPS), near_i.masked_fill(padded, -1)

    def _atom_embedding(self, atom_mask: torch.Tensor) -> torch.Tensor:
        B, M = atom_mask.shape
        kinds = (torch.arange(M, device=atom_mask.device) % NUM_MAIN_SEQ_ATOMS)[None, :].expand(B, -1)
        return self.graph_norm.compose(self.embedding(kinds) * atom_mask[:, :, None], atom_mask)

    def forward(self, atom_coords: torch.Tensor, atom_mask: torch.Tensor):
        """-> (atom_embedding [B,M,D], atom_cross_dists [B,M,K], atom_edge_index [B,M,K])."""
        dists, index = self._get_atom_graph(atom_coords, atom_mask)
        return self._atom_embedding(atom_mask), dists, index


class ResFeature(nn.Module):
    def __init__(self,
                 num_neighbours: int,
                 num_inside_dist_atoms: int = NUM_MAIN_SEQ_ATOMS,
                 num_inside_angle_atoms: int = NUM_MAIN_SEQ_ATOMS - 1,
                 num_inside_dihedral_atoms: int = NUM_MAIN_SEQ_ATOMS - 1,
                 num_cross_dist_atoms: int = NUM_MAIN_SEQ_ATOMS,
                 num_cross_angle_atoms: int = NUM_MAIN_SEQ_ATOMS - 1,
                 num_cross_dihedral_atoms: int = NUM_MAIN_SEQ_ATOMS - 1,
                 res_embedding_dim: int = DEFAULT_HIDDEN_DIM,
                 padding_len: int = 4500,
                 num_attn_layers: int = 2,
                 num_heads=4,
                 ffn_dim=DEFAULT_HIDDEN_DIM,
                 num_ffn_layers=2,
                 res_edge_embedding_dim: int = DEFAULT_HIDDEN_DIM,
                 num_edge_layers: int = 2,
                 dropout: float = 0.1):
        super().__init__()
        for name, v, lo in (("num_inside_dist_atoms", num_inside_dist_atoms, 2),
                            ("num_inside_angle_atoms", num_inside_angle_atoms, 3),
                            ("num_inside_dihedral_atoms", num_inside_dihedral_atoms, 4),
                            ("num_cross_dist_atoms", num_cross_dist_atoms, 2),
                            ("num_cross_angle_atoms", num_cross_angle_atoms, 3),
                            ("num_cross_dihedral_atoms", num_cross_dihedral_atoms, 4)):
            assert v >= lo, f"{name}({v}) must be at least {lo}."
        self.num_neighbours = num_neighbours
        self.num_inside_dist_atoms = num_inside_dist_atoms
        self.num_inside_angle_atoms = num_inside_angle_atoms
        self.num_inside_dihedral_atoms = num_inside_dihedral_atoms
        self.num_cross_dist_atoms = num_cross_dist_atoms
        self.num_cross_angle_atoms = num_cross_angle_atoms
        self.num_cross_dihedral_atoms = num_cross_dihedral_atoms
        self.raw_dim = (num_inside_dist_atoms * (num_inside_dist_atoms - 1) // 2 + num_inside_angle_atoms - 2
                        + num_inside_dihedral_atoms - 3)
        raw_edge_dim = (num_cross_dist_atoms ** 2 + (num_cross_angle_atoms - 1) ** 2
                        + (num_cross_dihedral_atoms - 2) ** 2)
        self.raw_project = nn.Linear(self.raw_dim, res_embedding_dim)
        self.res_embedding = RNABert(padding_len=padding_len, res_embedding_dim=res_embedding_dim,
                                     num_attn_layers=num_attn_layers, num_heads=num_heads, ffn_dim=ffn_dim,
                                     num_ffn_layers=num_ffn_layers, dropout=dropout)
        self.graph_norm = GraphNormalization(embedding_dim=res_embedding_dim)
        layers, width = [], raw_edge_dim
        for _ in range(num_edge_layers):
            layers += [nn.Linear(width, res_edge_embedding_dim), nn.GELU(), nn.Dropout(dropout)]
            width = res_edge_embedding_dim
        self.res_edge_embedding_layers = nn.Sequential(*layers)

    # -- pieces the reference exposes as methods --------------------------------------
    def _get_res_graph(self, coords: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """edge_index [B,N,num_neighbours] int64, -1 sentinels (reference: feature.py:205-256)."""
        _kernel_path(coords, mask)
        return ops.res_graph(coords, mask.to(torch.float32), self.num_neighbours)

    def _raw_features(self, coords, mask, edge_index):
        """(raw [B,N,raw_dim], raw_edge [B,N,K,raw_edge_dim]) (reference: feature.py:298-517, :531-535, :553-558)."""
        B, N = mask.shape
        raw, er = ops.res_raw_features(
            coords, mask.to(torch.float32), edge_index,
            (self.num_inside_dist_atoms, self.num_inside_angle_atoms, self.num_inside_dihedral_atoms),
            (self.num_cross_dist_atoms, self.num_cross_angle_atoms, self.num_cross_dihedral_atoms))
        return raw.reshape(B, N, -1), er.reshape(B, N, edge_index.shape[-1], -1)

    def _res_edge_embedding(self, coords, mask, edge_index, raw_edge=None):
        if raw_edge is None:
            _, raw_edge = self._raw_features(coords, mask, edge_index)
        dead = ((edge_index == -1) | (mask == 0).unsqueeze(-1)).unsqueeze(-1)
        if _kernel_path(*self.res_edge_embedding_layers.parameters(), training=self.training) \
                and not _dropout_active(self.res_edge_embedding_layers) and _widths_ok(self.res_edge_embedding_layers):
            g = ops.res_edge_rows(edge_index, mask.to(torch.float32))
            x = _flat(raw_edge)
            e = RNAMPNNEngine._mlp([(x, None)], _packed(self.res_edge_embedding_layers, last_bare=False), x.shape[0],
                                   row_scale=g.edge_keep)
            return e.reshape(*edge_index.shape, -1)
        return self.res_edge_embedding_layers(raw_edge).masked_fill(dead, 0.0)

    def _res_embedding(self, coords, mask, raw=None):
        if raw is None:
            raw, _ = ops.res_raw_features(
                coords, mask.to(torch.float32), None,
                (self.num_inside_dist_atoms, self.num_inside_angle_atoms, self.num_inside_dihedral_atoms),
                want_edges=False)
            raw = raw.reshape(*mask.shape, -1)
        if _kernel_path(*self.raw_project.parameters(), training=self.training) \
                and self.raw_project.out_features == ops.HIDDEN:
            x = _flat(raw)
            proj = RNAMPNNEngine._mlp([(x, None)], [(ops.pack_linear(self.raw_project.weight, self.raw_project.bias), False)],
                                      x.shape[0]).reshape(*mask.shape, -1)
        else:
            proj = self.raw_project(raw)
        return raw, self.res_embedding(proj, mask)

    def forward(self, coords: torch.Tensor, mask: torch.Tensor):
        """-> (raw [B,N,28], res_embedding [B,N,H], res_edge_embedding [B,N,K,H], edge_index [B,N,K] int64)."""
        _kernel_path(coords, mask)
        edge_index = self._get_res_graph(coords, mask)
        raw, raw_edge = self._raw_features(coords, mask, edge_index)
        res_edge_embedding = self._res_edge_embedding(coords, mask, edge_index, raw_edge)
        raw, res_embedding = self._res_embedding(coords, mask, raw)
        res_embedding = self.graph_norm(res_embedding, mask)
        return raw, res_embedding, res_edge_embedding, edge_index
