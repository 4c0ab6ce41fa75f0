"""Host-side data plumbing of the rnampnn package (out of the hot path, plain torch).  Mirrors the public
names the model code imports from rnampnn/utils/data.py: the collate ``featurize`` defines the padded-batch
layout the kernels consume (:110-142) and ``separate`` splits a concatenated per-residue vector (:594-604)."""
from __future__ import annotations

from typing import Dict, List, Tuple, Union

import torch

from ..config.glob import NUM_MAIN_SEQ_ATOMS, NUM_RES_TYPES


def featurize(batch: List[Dict[str, Union[str, torch.Tensor]]]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, List[str]]:
    """Pad a list of {'id', 'sequence' [L,4] one-hot, 'coordinates' [L,7,3]} items to the longest one:
    (sequences [B,N,4], coords [B,N,7,3], mask [B,N] prefix mask, ids)."""
    n = max(item['sequence'].shape[0] for item in batch)
    sequences = torch.zeros(len(batch), n, NUM_RES_TYPES, dtype=torch.float32)
    coords = torch.zeros(len(batch), n, NUM_MAIN_SEQ_ATOMS, 3, dtype=torch.float32)
    mask = torch.zeros(len(batch), n, dtype=torch.float32)
    ids = []
    for i, item in enumerate(batch):
        L = item['sequence'].shape[0]
        sequences[i, :L] = item['sequence']
        coords[i, :L] = item['coordinates']
        mask[i, :L] = 1
        ids.append(item['id'])
    return sequences, coords, mask, ids


def separate(concat: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """[sum L] -> zero-padded [B, max L] by the per-structure lengths."""
    lens = [int(v) for v in lengths.tolist()]
    out = torch.zeros(len(lens), max(lens) if lens else 0, dtype=concat.dtype, device=concat.device)
    at = 0
    for b, L in enumerate(lens):
        out[b, :L] = concat[at:at + L]
        at += L
    return out
