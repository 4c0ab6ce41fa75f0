"""Drop-in ``rnampnn.model.functional``: GraphNormalization, Readout, RNABert, RawFFN
(reference: rnampnn/model/functional.py:7-202).

Same constructor arguments, parameter names and forward signatures as the reference.  On CUDA tensors
without autograd the forward runs on the sm_100a kernels through the C ABI (``rmpnn_b200.ops``): a
two-launch GraphNorm, self-attention over the VALID rows only (the reference pads every structure to
``padding_len`` = 4500 rows and masks the padding), and the fused row-MLP for every Linear stack.  With
autograd enabled (``train()`` mode) the modules compose the same math from differentiable torch-CUDA ops --
the interim backward path.  CPU tensors are refused; widths the kernels are not built for (hidden != 128)
use the torch-CUDA composition.
"""
from __future__ import annotations

import torch
from torch import nn

from rmpnn_b200 import ops
from rmpnn_b200.engine_rnampnn import BertWeights, RNAMPNNEngine
from ..config.glob import NUM_RES_TYPES, SEPS


def _kernel_path(*tensors, training: bool = True) -> bool:
    """True -> C-ABI kernels; False -> differentiable torch-CUDA composition.  Same rule as the rdesign mirror:
    train() + autograd -> composition; eval() -> kernels, unless a non-parameter input requires grad or
    ``RMPNN_EVAL_AUTOGRAD=1`` asks for gradients through an eval-mode model."""
    import os
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("rnampnn (B200 build) runs on CUDA tensors only; there is no CPU path")
    if not torch.is_grad_enabled():
        return True
    if training or os.environ.get("RMPNN_EVAL_AUTOGRAD", "0") == "1":
        return not any(t.requires_grad for t in tensors)
    return not any(t.requires_grad and not isinstance(t, nn.Parameter) for t in tensors)


def _dropout_active(module: nn.Module) -> bool:
    return module.training and any(isinstance(m, nn.Dropout) and m.p > 0 for m in module.modules())


def _linears(seq):
    return [m for m in seq if isinstance(m, nn.Linear)]


def _packed(seq, last_bare: bool):
    lin = _linears(seq)
    return [(ops.pack_linear(m.weight, m.bias), not (last_bare and n == len(lin) - 1)) for n, m in enumerate(lin)]


def _widths_ok(seq) -> bool:
    return all(m.out_features % 128 == 0 for m in _linears(seq))


def _flat(x):
    return x.detach().to(torch.float32).reshape(-1, x.shape[-1]).contiguous()


class GraphNormalization(nn.Module):
    """Per-graph, per-channel normalisation over the valid rows; the padded rows the tensor holds add
    ``mean^2`` each to the variance (reference: functional.py:33-38)."""

    def __init__(self, embedding_dim: int):
        super().__init__()
        self.scale = nn.Parameter(torch.ones(1, 1, embedding_dim))
        self.shift = nn.Parameter(torch.zeros(1, 1, embedding_dim))

    def forward(self, features: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        if _kernel_path(features, mask, self.scale, self.shift, training=self.training) \
                and features.shape[-1] == ops.HIDDEN and features.dim() == 3:
            B, N, _ = features.shape
            out = ops.graph_norm(_flat(features), None, mask.to(torch.float32).reshape(-1).contiguous(), B, N,
                                 self.scale.detach().reshape(-1).contiguous(), self.shift.detach().reshape(-1).contiguous())
            return out.reshape(features.shape)
        return self.compose(features, mask)

    def compose(self, features: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """The same normalisation from differentiable torch ops (any device): the train()-mode path, and what the
        plain-PyTorch passthrough modules (AtomFeature, AtomMPNN) call."""
        m = mask.unsqueeze(-1).to(features.dtype)
        xm = features * m
        n = m.sum(dim=1, keepdim=True)
        n = torch.where(n == 0, torch.ones_like(n), n)
        mean = xm.sum(dim=1, keepdim=True) / n
        var = ((xm - mean) ** 2).sum(dim=1, keepdim=True) / n
        return ((features - mean) / torch.sqrt(var + SEPS) * self.scale + self.shift) * m


class Readout(nn.Module):
    """(num_layers-1) x [Linear, GELU, Dropout] then Linear -> 4 logits, times the mask."""

    def __init__(self, embedding_dim: int, readout_hidden_dim: int, num_layers: int, dropout: float = 0.1):
        super().__init__()
        layers, width = [], embedding_dim
        for _ in range(num_layers - 1):
            layers += [nn.Linear(width, readout_hidden_dim), nn.GELU(), nn.Dropout(dropout)]
            width = readout_hidden_dim
        layers.append(nn.Linear(width, NUM_RES_TYPES))
        self.readout_layers = nn.Sequential(*layers)

    def forward(self, res_embedding: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        lin = _linears(self.readout_layers)
        if not _kernel_path(res_embedding, mask, *self.parameters(), training=self.training) or _dropout_active(self) \
                or not all(m.out_features % 128 == 0 for m in lin[:-1]):
            return self.readout_layers(res_embedding) * mask.unsqueeze(-1)
        x = _flat(res_embedding)
        if len(lin) > 1:
            x = RNAMPNNEngine._mlp([(x, None)], [(ops.pack_linear(m.weight, m.bias), True) for m in lin[:-1]], x.shape[0])
        logits, _, _ = ops.readout_head_masked(x, lin[-1].weight.detach().contiguous(), lin[-1].bias.detach().contiguous(),
                                               mask.to(torch.float32).reshape(-1).contiguous(), want_probs=False,
                                               want_samples=False)
        return logits.reshape(*res_embedding.shape[:-1], NUM_RES_TYPES)


class RNABert(nn.Module):
    """num_attn_layers x [x += MHA(x); GraphNorm over a padding_len-row padding] then an FFN, masked."""

    def __init__(self, padding_len: int, res_embedding_dim: int, num_attn_layers: int, num_heads: int, ffn_dim: int,
                 num_ffn_layers: int, dropout: float = 0.1):
        super().__init__()
        self.padding_len = padding_len
        self.num_heads = num_heads
        self.bi_attention_layers = nn.ModuleList([
            nn.MultiheadAttention(embed_dim=res_embedding_dim, num_heads=num_heads, dropout=dropout, batch_first=True)
            for _ in range(num_attn_layers)])
        self.graph_norm_layers = nn.ModuleList([GraphNormalization(res_embedding_dim) for _ in range(num_attn_layers)])
        layers, width = [], res_embedding_dim
        for _ in range(num_ffn_layers):
            layers += [nn.Linear(width, ffn_dim), nn.GELU(), nn.Dropout(dropout)]
            width = ffn_dim
        layers.append(nn.Linear(ffn_dim, res_embedding_dim))
        self.ffn_layers = nn.Sequential(*layers)

    def _bert_weights(self) -> BertWeights:
        bw = BertWeights()
        for att, gn in zip(self.bi_attention_layers, self.graph_norm_layers):
            bw.attn.append((ops.pack_linear(att.in_proj_weight, att.in_proj_bias),
                            ops.pack_linear(att.out_proj.weight, att.out_proj.bias)))
            bw.norms.append((gn.scale.detach().reshape(-1).contiguous(), gn.shift.detach().reshape(-1).contiguous()))
        bw.ffn = _packed(self.ffn_layers, last_bare=True)
        return bw

    def _kernel_ok(self, x) -> bool:
        E = x.shape[-1]
        hd = E // self.num_heads
        return E == ops.HIDDEN and _widths_ok(self.ffn_layers) and (len(self.bi_attention_layers) == 0 or hd in (8, 16, 32))

    def forward(self, res_embedding: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        if _kernel_path(res_embedding, mask, *self.parameters(), training=self.training) and not _dropout_active(self) \
                and self._kernel_ok(res_embedding):
            B, N, _ = res_embedding.shape
            eng = RNAMPNNEngine.__new__(RNAMPNNEngine)
            eng.padding_len = self.padding_len
            out = eng.bert(self._bert_weights(), _flat(res_embedding), mask.to(torch.float32).reshape(-1).contiguous(),
                           B, N, self.num_heads)
            return out.reshape(res_embedding.shape)
        # differentiable composition on the N real rows: padded keys carry no weight, padded queries are masked,
        # and the (padding_len - N) extra rows of the reference only enter GraphNorm's variance
        x = res_embedding
        m = mask.unsqueeze(-1).to(x.dtype)
        N = x.shape[1]
        for att, gn in zip(self.bi_attention_layers, self.graph_norm_layers):
            x = x + att(query=x, key=x, value=x, key_padding_mask=~mask.bool(), need_weights=False)[0]
            xm = x * m
            n = m.sum(dim=1, keepdim=True)
            n = torch.where(n == 0, torch.ones_like(n), n)
            mean = xm.sum(dim=1, keepdim=True) / n
            var = (((xm - mean) ** 2).sum(dim=1, keepdim=True) + (self.padding_len - N) * mean ** 2) / n
            x = ((x - mean) / torch.sqrt(var + SEPS) * gn.scale + gn.shift) * m
        return self.ffn_layers(x) * m


class RawFFN(nn.Module):
    """raw (28) -> FFN -> raw_embedding_dim, then GraphNorm."""

    def __init__(self, raw_dim: int, num_raw_ffn_dim: int, num_raw_ffn_layers: int, raw_embedding_dim: int,
                 dropout: float = 0.1):
        super().__init__()
        layers, width = [], raw_dim
        for _ in range(num_raw_ffn_layers):
            layers += [nn.Linear(width, num_raw_ffn_dim), nn.GELU(), nn.Dropout(dropout)]
            width = num_raw_ffn_dim
        layers.append(nn.Linear(num_raw_ffn_dim, raw_embedding_dim))
        self.raw_ffn = nn.Sequential(*layers)
        self.graph_norm = GraphNormalization(raw_embedding_dim)

    def forward(self, raw: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        if _kernel_path(raw, mask, *self.parameters(), training=self.training) and not _dropout_active(self) \
                and _widths_ok(self.raw_ffn) and _linears(self.raw_ffn)[-1].out_features == ops.HIDDEN:
            x = _flat(raw)
            y = RNAMPNNEngine._mlp([(x, None)], _packed(self.raw_ffn, last_bare=True), x.shape[0])
            return self.graph_norm(y.reshape(*raw.shape[:-1], -1), mask)
        return self.graph_norm(self.raw_ffn(raw), mask)
