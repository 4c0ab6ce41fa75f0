"""Drop-in ``rdesign.model.functional``: Normalize, Readout, scatter_sum.

Same constructor arguments, parameter names and forward signatures as the reference
(rdesign/model/functional.py:51-127); on CUDA tensors without autograd the forward
runs on the sm_100a kernels through the C ABI.  With autograd enabled the modules
compose the same math from differentiable torch-CUDA ops (the interim backward path
until ``_bwd`` kernels land).  CPU tensors are refused.

``GraphNormalization`` and ``RNABert`` of the reference file are dead code for
``RNAModel`` (never instantiated) and are not provided.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from rmpnn_b200 import ops


def _kernel_path(*tensors, training: bool = True) -> bool:
    """True -> run the C-ABI kernels; False -> differentiable torch-CUDA composition.

    ``train()`` mode with autograd enabled and something to differentiate takes the composition.  In ``eval()``
    mode the kernels run (their outputs carry no autograd graph) unless autograd is on and either a non-parameter
    input requires grad (saliency / gradient checks through an eval-mode model) or ``RMPNN_EVAL_AUTOGRAD=1`` is set
    (fine-tuning with dropout off).  Parameters alone do not switch an eval-mode model to the composition: the
    reference's ``predict()`` (rdesign/model/rdesign.py:144-172) runs without ``torch.no_grad()`` and must stay on
    the kernels."""
    import os
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("rdesign (B200 build) runs on CUDA tensors only; there is no CPU path")
    if not torch.is_grad_enabled():
        return True
    if training or os.environ.get("RMPNN_EVAL_AUTOGRAD", "0") == "1":
        return not any(t.requires_grad for t in tensors)
    return not any(t.requires_grad and not isinstance(t, nn.Parameter) for t in tensors)


def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out: Optional[torch.Tensor] = None,
                dim_size: Optional[int] = None) -> torch.Tensor:
    """Sum rows of ``src`` that share an index (reference: functional.py:51-66)."""
    if dim < 0:
        dim += src.dim()
    if out is None:
        size = list(src.shape)
        size[dim] = dim_size if dim_size is not None else (0 if index.numel() == 0 else int(index.max()) + 1)
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.index_add_(dim, index, src)


class Normalize(nn.Module):
    """gain * (x - mean) / (sqrt(var_unbiased + eps) + eps) + bias over the last dim."""

    def __init__(self, features, epsilon=1e-6):
        super().__init__()
        self.gain = nn.Parameter(torch.ones(features))
        self.bias = nn.Parameter(torch.zeros(features))
        self.epsilon = epsilon

    def forward(self, x, dim=-1):
        use_kernel = _kernel_path(x, self.gain, self.bias, training=self.training)
        if use_kernel and dim in (-1, x.dim() - 1) and x.shape[-1] == ops.HIDDEN and x.dtype == torch.float32:
            flat = x.reshape(-1, ops.HIDDEN).contiguous()
            out = ops.add_layernorm(flat, None, self.gain.detach(), self.bias.detach(), self.epsilon,
                                    reference_normalize=True)
            return out.reshape(x.shape)
        mu = x.mean(dim, keepdim=True)
        sigma = torch.sqrt(x.var(dim, keepdim=True) + self.epsilon)
        shape = [1] * x.dim()
        shape[dim] = -1
        return self.gain.view(shape) * (x - mu) / (sigma + self.epsilon) + self.bias.view(shape)


class Readout(nn.Module):
    """(num_layers-1) x [Linear, GELU, Dropout] then Linear -> 4 logits."""

    def __init__(self, embedding_dim: int, readout_hidden_dim: int, num_layers: int, dropout: float = 0.1):
        super().__init__()
        layers, width = [], embedding_dim
        for _ in range(num_layers - 1):
            layers += [nn.Linear(width, readout_hidden_dim), nn.GELU(), nn.Dropout(dropout)]
            width = readout_hidden_dim
        layers.append(nn.Linear(width, 4))
        self.readout_layers = nn.Sequential(*layers)

    def _linears(self):
        return [m for m in self.readout_layers if isinstance(m, nn.Linear)]

    def _dropout_active(self):
        return self.training and any(isinstance(m, nn.Dropout) and m.p > 0 for m in self.readout_layers)

    def forward(self, res_embedding: torch.Tensor) -> torch.Tensor:
        if not _kernel_path(res_embedding, *self.parameters(), training=self.training) or self._dropout_active():
            return self.readout_layers(res_embedding)
        lin = self._linears()
        x = res_embedding.detach().to(torch.float32)
        lead = x.shape[:-1]
        x = x.reshape(-1, x.shape[-1]).contiguous()
        for m in lin[:-1]:
            y = torch.empty(x.shape[0], m.out_features, dtype=torch.float32, device=x.device)
            ops.row_mlp([(x, None)], [(ops.pack_linear(m.weight, m.bias), True)], ops.EPI_STORE, y, x.shape[0])
            x = y
        logits, _, _ = ops.readout_head(x, lin[-1].weight.detach().contiguous(), lin[-1].bias.detach().contiguous(),
                                        want_probs=False, want_samples=False)
        return logits.reshape(*lead, 4)
