"""Drop-in ``rnampnn.model.pool`` (reference: rnampnn/model/pool.py:7-60).  ``AtomPool`` is not imported by any live
path of the reference; it is kept as a plain-PyTorch passthrough so the name resolves."""
import torch
from torch import nn
from torch.nn import functional as F

from ..config.glob import NUM_MAIN_SEQ_ATOMS


class AtomPool(nn.Module):
    """Softmax-weighted pooling of a residue's seven atom embeddings; the weights come from an MLP on the residue's
    raw features (masked logits are set to 0 before the softmax, as in the reference)."""

    def __init__(self, raw_dim: int, atom_pool_hidden_dim: int, num_layers: int, dropout: float):
        super().__init__()
        self.num_layers = num_layers
        self.atom_pool_hidden_dim = atom_pool_hidden_dim
        layers, width = [], raw_dim
        for _ in range(num_layers - 1):
            layers += [nn.Linear(width, atom_pool_hidden_dim), nn.GELU(), nn.Dropout(dropout)]
            width = atom_pool_hidden_dim
        layers.append(nn.Linear(atom_pool_hidden_dim, NUM_MAIN_SEQ_ATOMS))
        self.weight_layers = nn.Sequential(*layers)

    def forward(self, atom_embedding: torch.Tensor, atom_mask: torch.Tensor, raw: torch.Tensor) -> torch.Tensor:
        """atom_embedding [B,N*7,D], atom_mask [B,N*7], raw [B,N,raw_dim] -> [B,N,D]."""
        B, N, _ = raw.shape
        per_residue = atom_embedding.reshape(B, N, NUM_MAIN_SEQ_ATOMS, -1)
        w = F.softmax(self.weight_layers(raw) * atom_mask.reshape(B, N, NUM_MAIN_SEQ_ATOMS), dim=-1)
        return (per_residue * w[..., None]).sum(dim=2)


class ResPool(nn.Module):
    """Stub in the reference as well (forward returns None)."""

    def forward(self, res_embedding: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        return None
