"""Drop-in ``rdesign.model.mpnn.MPNNLayer`` (reference: rdesign/model/mpnn.py:5-38).

Standalone use keeps the reference's calling convention -- the caller passes the
already concatenated ``h_EV = [h_E | h_V[centre] | h_V[neighbour]]`` -- and runs three
launches: fused message MLP + segmented sum, add + LayerNorm, fused FFN + residual +
LayerNorm.  ``RNAModel.forward`` does not go through here: it owns ``h_E``/``h_V``/
``E_idx`` and lets the kernel gather ``h_V`` rows itself, so ``h_EV`` never exists in HBM.
"""
from __future__ import annotations

import torch
from torch import nn

from rmpnn_b200 import ops
from .functional import _kernel_path, scatter_sum


def _gelu_stack(widths_in, widths_out, dropout):
    mods = []
    for i, o in zip(widths_in, widths_out):
        mods += [nn.Linear(i, o), nn.GELU(), nn.Dropout(dropout)]
    return mods


class MPNNLayer(nn.Module):
    def __init__(self, num_hidden, num_in, num_message_layers, num_dense_layers, dim_dense_layers, dropout=0.1,
                 scale=30):
        super().__init__()
        self.num_hidden, self.num_in, self.scale = num_hidden, num_in, scale
        self.norm1 = nn.LayerNorm(num_hidden)
        self.norm2 = nn.LayerNorm(num_hidden)
        msg_in = [num_hidden + num_in] + [num_hidden] * (num_message_layers - 1)
        self.message_layers = nn.Sequential(*_gelu_stack(msg_in, [num_hidden] * num_message_layers, dropout))
        # the reference carries the running width from the message loop into the dense loop
        first = num_hidden if num_message_layers > 0 else num_hidden + num_in
        den_in = [first] + [dim_dense_layers] * (num_dense_layers - 1)
        last_in = dim_dense_layers if num_dense_layers > 0 else first
        self.dense = nn.Sequential(*_gelu_stack(den_in, [dim_dense_layers] * num_dense_layers, dropout),
                                   nn.Linear(last_in, num_hidden))

    @staticmethod
    def _linears(seq):
        return [m for m in seq if isinstance(m, nn.Linear)]

    def _dropout_active(self):
        return self.training and any(isinstance(m, nn.Dropout) and m.p > 0 for m in self.modules())

    def forward(self, h_V, h_E, edge_idx) -> torch.Tensor:
        """h_V [Nv,H]; h_E = concatenated edge input [Ne, H+num_in]; edge_idx [2,Ne] (row 0 = centre)."""
        if not _kernel_path(h_V, h_E, *self.parameters(), training=self.training) or self._dropout_active() \
                or self.num_hidden != ops.HIDDEN:
            dh = scatter_sum(self.message_layers(h_E), edge_idx[0], dim=0, dim_size=h_V.shape[0]) / self.scale
            h_V = self.norm1(h_V + dh)
            return self.norm2(h_V + self.dense(h_V))
        h_V = h_V.detach().to(torch.float32).contiguous()
        h_E = h_E.detach().to(torch.float32).contiguous()
        centre = edge_idx[0].to(torch.int32).contiguous()
        msg = [(ops.pack_linear(m.weight, m.bias), True) for m in self._linears(self.message_layers)]
        den = self._linears(self.dense)
        den = [(ops.pack_linear(m.weight, m.bias), True) for m in den[:-1]] + \
              [(ops.pack_linear(den[-1].weight, den[-1].bias), False)]
        agg = torch.zeros_like(h_V)
        ops.row_mlp([(h_E, None)], msg, ops.EPI_SEGSUM, agg, h_E.shape[0], seg_id=centre, alpha=1.0 / self.scale)
        x1 = ops.add_layernorm(h_V, agg, self.norm1.weight.detach(), self.norm1.bias.detach(), self.norm1.eps)
        out = torch.empty_like(h_V)
        ops.row_mlp([(x1, None)], den, ops.EPI_RESID_LN, out, h_V.shape[0], gain=self.norm2.weight.detach(),
                    beta=self.norm2.bias.detach(), resid=x1)
        return out
