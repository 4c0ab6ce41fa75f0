"""Drop-in ``rnampnn.model.mpnn.ResMPNN`` (reference: rnampnn/model/mpnn.py:115-294).

``forward`` runs three launches on the sm_100a kernels: the fused message MLP (the kernel gathers
``h[centre] | h[neighbour] | e`` itself, applies the two Linear+GELU stages and accumulates the masked mean per
centre), the two-launch GraphNorm with the residual add fused in, and the fused edge MLP with its in-place
``e +=`` epilogue.  ``AtomMPNN`` (reference :7-112) is never reached from ``RNAMPNN.forward``; it is kept as a
plain-PyTorch passthrough so that ``test.py::test_atom_mpnn`` and user code importing it keep working.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import nn

from rmpnn_b200 import ops
from .functional import GraphNormalization, _dropout_active, _flat, _kernel_path, _packed


class AtomMPNN(nn.Module):
    """Atom-level message passing (passthrough, plain PyTorch): message MLP on [h_neighbour | h_atom | distance],
    mean over the valid neighbour slots, residual, GraphNormalization."""

    def __init__(self, atom_embedding_dim: int, depth_atom_mpnn: int, dropout: float = 0.1):
        super().__init__()
        self.graph_norm = GraphNormalization(embedding_dim=atom_embedding_dim)
        layers, width = [], 2 * atom_embedding_dim + 1
        for _ in range(depth_atom_mpnn):
            layers += [nn.Linear(width, atom_embedding_dim), nn.GELU(), nn.Dropout(dropout)]
            width = atom_embedding_dim
        self.message_layers = nn.Sequential(*layers)

    def message(self, atom_embedding, atom_cross_dists, atom_edge_index, atom_mask):
        """[B,M,K,D] messages; slots holding -1 are zero."""
        B, M, K = atom_edge_index.shape
        h = atom_embedding * atom_mask[:, :, None]
        rows = atom_edge_index.clamp(min=0)                    # -1 reads atom 0; the slot is zeroed below
        nbr = torch.gather(h, 1, rows.reshape(B, M * K, 1).expand(-1, -1, h.shape[-1])).reshape(B, M, K, -1)
        pair = torch.cat([nbr, h[:, :, None, :].expand(-1, -1, K, -1), atom_cross_dists[..., None]], dim=-1)
        return self.message_layers(pair) * (atom_edge_index != -1)[..., None].float()

    @staticmethod
    def aggregation(atom_embedding, messages, atom_edge_index, atom_mask):
        count = (atom_edge_index != -1).sum(dim=-1, keepdim=True).float().clamp(min=1.0)
        return (atom_embedding + messages.sum(dim=2) / count) * atom_mask[:, :, None]

    def forward(self, atom_embedding, atom_cross_dists, atom_edge_index, atom_mask):
        messages = self.message(atom_embedding, atom_cross_dists, atom_edge_index, atom_mask)
        return self.graph_norm.compose(self.aggregation(atom_embedding, messages, atom_edge_index, atom_mask), atom_mask)


class ResMPNN(nn.Module):
    def __init__(self, res_embedding_dim: int, res_edge_embedding_dim: int, depth_res_mpnn: int, num_edge_layers: int,
                 dropout: float):
        super().__init__()
        self.graph_norm = GraphNormalization(embedding_dim=res_embedding_dim)
        layers, width = [], res_embedding_dim * 2 + res_edge_embedding_dim
        for _ in range(depth_res_mpnn):
            layers += [nn.Linear(width, res_embedding_dim), nn.GELU(), nn.Dropout(dropout)]
            width = res_embedding_dim
        self.message_layers = nn.Sequential(*layers)
        layers, width = [], res_embedding_dim * 2 + res_edge_embedding_dim
        for _ in range(num_edge_layers):
            layers += [nn.Linear(width, res_edge_embedding_dim), nn.GELU(), nn.Dropout(dropout)]
            width = res_edge_embedding_dim
        self.edge_layers = nn.Sequential(*layers)

    def _kernel_ok(self, h, e, mask) -> bool:
        return (_kernel_path(h, e, mask, *self.parameters(), training=self.training) and not _dropout_active(self)
                and h.shape[-1] == ops.HIDDEN and e.shape[-1] == ops.HIDDEN
                and len(self.message_layers) > 0 and len(self.edge_layers) > 0)

    @staticmethod
    def _gather(h, edge_index):
        safe = edge_index.clamp(min=0)
        B, N, K = safe.shape
        return torch.gather(h, 1, safe.reshape(B, N * K, 1).expand(-1, -1, h.shape[-1])).reshape(B, N, K, -1)

    def _inputs(self, h, e, edge_index):
        K = edge_index.shape[-1]
        return torch.cat([h.unsqueeze(2).expand(-1, -1, K, -1), self._gather(h, edge_index), e], dim=-1)

    def message(self, res_embedding, res_edge_embedding, edge_index, mask) -> torch.Tensor:
        """[B,N,K,H] messages, zero on -1 edges (reference: mpnn.py:154-194)."""
        h = res_embedding * mask.unsqueeze(-1)
        valid = (edge_index != -1).unsqueeze(-1).to(h.dtype)
        if self._kernel_ok(res_embedding, res_edge_embedding, mask):
            g = ops.res_edge_rows(edge_index, mask.to(torch.float32))
            hf, ef = _flat(h), _flat(res_edge_embedding)
            out = torch.empty(ef.shape[0], ops.HIDDEN, dtype=torch.float32, device=hf.device)
            ops.row_mlp([(hf, g.ctr_row), (hf, g.nbr_row), (ef, None)], _packed(self.message_layers, False),
                        ops.EPI_STORE, out, ef.shape[0], row_scale=valid.reshape(-1).to(torch.float32).contiguous())
            return out.reshape(res_edge_embedding.shape[:-1] + (ops.HIDDEN,))
        return self.message_layers(self._inputs(h, res_edge_embedding, edge_index)) * valid

    @staticmethod
    def aggregation(mask, res_embedding, messages, edge_index) -> torch.Tensor:
        """(h + mean over valid neighbours of the messages) * mask (reference: mpnn.py:196-227)."""
        cnt = (edge_index != -1).sum(dim=-1, keepdim=True).to(messages.dtype)
        cnt = torch.where(cnt == 0, torch.ones_like(cnt), cnt)
        return (res_embedding + messages.sum(dim=2) / cnt) * mask.unsqueeze(-1)

    def _update_edges(self, res_embedding, res_edge_embedding, edge_index) -> torch.Tensor:
        res_edge_embedding += self.edge_layers(self._inputs(res_embedding, res_edge_embedding, edge_index))
        return res_edge_embedding

    def forward(self, res_embedding, res_edge_embedding, edge_index, mask) -> Tuple[torch.Tensor, torch.Tensor]:
        if not self._kernel_ok(res_embedding, res_edge_embedding, mask):
            messages = self.message(res_embedding, res_edge_embedding, edge_index, mask)
            res_embedding = self.aggregation(mask, res_embedding, messages, edge_index)
            res_embedding = self.graph_norm(res_embedding, mask)
            return res_embedding, self._update_edges(res_embedding, res_edge_embedding, edge_index)
        B, N, K = edge_index.shape
        mask_flat = mask.to(torch.float32).reshape(-1).contiguous()
        g = ops.res_edge_rows(edge_index, mask_flat.reshape(B, N))
        h = _flat(res_embedding)
        hm = ops.scale_rows(h.clone(), mask_flat)                 # the message reads h * mask (mpnn.py:170)
        e = res_edge_embedding.detach()
        e_flat = e.reshape(-1, ops.HIDDEN) if (e.dtype == torch.float32 and e.is_contiguous()) else _flat(e)
        agg = torch.zeros_like(h)
        seg = [(hm, g.ctr_row), (hm, g.nbr_row), (e_flat, None)]
        ops.row_mlp(seg, _packed(self.message_layers, False), ops.EPI_SEGSUM, agg, e_flat.shape[0], seg_id=g.ctr_row,
                    alpha=1.0, row_scale=g.edge_w)
        h = ops.graph_norm(h, agg, mask_flat, B, N, self.graph_norm.scale.detach().reshape(-1).contiguous(),
                           self.graph_norm.shift.detach().reshape(-1).contiguous(), zero_b=True)
        ops.row_mlp([(h, g.ctr_row), (h, g.nbr_row), (e_flat, None)], _packed(self.edge_layers, False),
                    ops.EPI_RESID_ADD, e_flat, e_flat.shape[0], resid=e_flat)        # in place, like the reference
        return h.reshape(res_embedding.shape), e_flat.reshape(res_edge_embedding.shape)
