"""Drop-in ``rnampnn.utils.data`` (reference: rnampnn/utils/data.py): the in-memory dataset built from
``coords/*.npy`` + ``seqs/*.fasta``, its augmentations, the collate ``featurize`` that defines the padded-batch
layout the kernels consume (:110-142), ``RNADataModule`` (:397-539) and ``separate`` (:594-604).

Host I/O, out of the B200 hot path, plain numpy / torch.  FASTA files are read with a few lines of Python, so
biopython is not needed; ``RNADataModule`` derives from Lightning's data module when Lightning is installed and
from ``object`` otherwise (same methods either way).  Plot / CSV report helpers of the reference file
(``analyse_dataset``, ``draw_recovery_scatter``) are analysis scripts and are not provided."""
from __future__ import annotations

import os
import random
from typing import Dict, List, Sequence, Tuple, Union

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from ..config.glob import DATA_PATH, MIN_LEN, NUM_MAIN_SEQ_ATOMS, NUM_RES_TYPES, VOCAB

try:
    from pytorch_lightning import LightningDataModule as _DataModuleBase
except ImportError:
    _DataModuleBase = object

Item = Dict[str, Union[str, torch.Tensor]]


def read_fasta(path: str) -> List[Tuple[str, str]]:
    """[(record id, sequence)] of a FASTA file (id = first word of the header line)."""
    records, name, chunks = [], None, []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith('>'):
                if name is not None:
                    records.append((name, ''.join(chunks)))
                name, chunks = (line[1:].split() or [''])[0], []
            elif line and name is not None:
                chunks.append(line)
    if name is not None:
        records.append((name, ''.join(chunks)))
    return records


def one_hot(sequence: str) -> torch.Tensor:
    """[L,4] one-hot over A,U,C,G; any other letter stays an all-zero row (data.py:224-238)."""
    out = torch.zeros(len(sequence), NUM_RES_TYPES, dtype=torch.float32)
    for i, ch in enumerate(sequence):
        if ch in VOCAB:
            out[i, VOCAB[ch]] = 1.0
    return out


def fill_missing_atoms(coords: np.ndarray, rng=np.random) -> np.ndarray:
    """Missing (NaN) atoms are parked near a present atom of the same residue: backbone atoms (columns 0-5) at
    1.5 A in a random direction from the residue's first present atom, the base atom (column 6) at 4 A from
    column 5 when that one is present; whatever is still missing becomes 0 (data.py:189-222).  Columns are handled
    in order, so an atom filled earlier can serve as the anchor of a later one, as in the reference."""
    c = np.array(coords, dtype=np.float32, copy=True)

    def unit():
        v = rng.randn(3)
        return v / np.linalg.norm(v)

    for a in range(c.shape[1]):
        for r in np.flatnonzero(np.isnan(c[:, a, :]).any(axis=1)):
            if a < 6:
                present = np.flatnonzero(~np.isnan(c[r]).any(axis=1))
                if present.size:
                    c[r, a] = c[r, present[0]] + 1.5 * unit()
            elif a == 6 and not np.isnan(c[r, 5]).any():
                c[r, a] = c[r, 5] + NUM_RES_TYPES * unit()
    return np.nan_to_num(c, nan=0.0)


def featurize(batch: Sequence[Item]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, List[str]]:
    """Pad {'id', 'sequence' [L,4], 'coordinates' [L,7,3]} items to the longest one:
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


class RNADataset(Dataset):
    """List of {'id', 'sequence', 'coordinates'} held in memory."""

    def __init__(self, data: List[Item]) -> None:
        super().__init__()
        self.data = data

    @classmethod
    def from_path(cls, data_path: str, is_predict: bool = False) -> "RNADataset":
        """``<data_path>/coords/<id>.npy`` [L,7,3] (+ ``seqs/<id>.fasta`` unless ``is_predict``, which stands in an
        all-'A' sequence of the right length)."""
        coords_dir, seqs_dir = os.path.join(data_path, "coords"), os.path.join(data_path, "seqs")
        data = []
        for fname in os.listdir(coords_dir):
            if not fname.endswith(".npy"):
                continue
            rna_id = os.path.splitext(fname)[0]
            c = np.load(os.path.join(coords_dir, fname)).astype(np.float32)
            coordinates = torch.from_numpy(fill_missing_atoms(c) if np.isnan(c).any() else c)
            sequence = one_hot("A" * c.shape[0]) if is_predict else cls._load_sequence(seqs_dir, rna_id)
            data.append({'id': rna_id, 'sequence': sequence, 'coordinates': coordinates})
        return cls(data)

    _one_hot_encode = staticmethod(one_hot)

    @staticmethod
    def _load_sequence(seqs_path: str, rna_id: str) -> torch.Tensor:
        path = os.path.join(seqs_path, f"{rna_id}.fasta")
        if not os.path.exists(path):
            raise FileNotFoundError(f"Sequence file for RNA ID {rna_id} not found.")
        records = read_fasta(path)
        if not records:
            raise ValueError(f"No valid sequence found in {path}.")
        return one_hot(records[0][1])

    # ---- filters and augmentations (data.py:258-324)
    def filter_by_min_length(self, min_len: int) -> None:
        self.data = [d for d in self.data if d['sequence'].shape[0] >= min_len]

    def filter_by_max_length(self, max_len: int) -> None:
        self.data = [d for d in self.data if d['sequence'].shape[0] <= max_len]

    def noise_augmentation(self, num_gen: int) -> None:
        """``num_gen`` copies of random items with N(0, 0.01^2) noise on every coordinate."""
        for _ in range(num_gen):
            src = self.data[np.random.randint(len(self.data))]
            noisy = src['coordinates'] + torch.normal(mean=0, std=1e-2, size=src['coordinates'].shape)
            self.data.append({'id': src['id'], 'sequence': src['sequence'], 'coordinates': noisy})

    def slice_augmentation(self, num_gen: int, min_len: int = MIN_LEN) -> None:
        """``num_gen`` random windows of ``min_len`` residues cut from items longer than that."""
        for _ in range(num_gen):
            pool = [d for d in self.data if d['sequence'].shape[0] > min_len]
            if not pool:
                raise ValueError("No sequences longer than min_len available for slicing.")
            src = random.choice(pool)
            at = random.randint(0, src['sequence'].shape[0] - min_len)
            self.data.append({'id': src['id'], 'sequence': src['sequence'][at:at + min_len],
                              'coordinates': src['coordinates'][at:at + min_len]})

    # ---- access
    def getitem_by_key(self, target_id: str) -> List[Item]:
        return [d for d in self.data if d['id'] == target_id]

    def get_ids(self) -> List[str]:
        return list({d['id'] for d in self.data})

    def shuffle(self, return_perm: bool = False):
        perm = torch.randperm(len(self.data))
        self.data = [self.data[i] for i in perm]
        return (self, perm) if return_perm else self

    def __len__(self) -> int:
        return len(self.data)

    len = __len__

    def __getitem__(self, idx: int) -> Item:
        return self.data[idx]


class RNADataModule(_DataModuleBase):
    """train / val split by RNA id (augmented copies of a structure stay on one side), loaders with ``featurize``."""
    num_workers = 8

    def __init__(self, data_path=DATA_PATH, split_ratio: Union[float, List[float]] = 0.9, batch_size: int = 8,
                 noise_augmentation: Union[int, None] = None, slice_augmentation: Union[int, None] = None,
                 min_len: Union[int, None] = None, max_len: Union[int, None] = None) -> None:
        super().__init__()
        self.data_path, self.split_ratio, self.batch_size = data_path, split_ratio, batch_size
        self.noise_augmentation, self.slice_augmentation = noise_augmentation, slice_augmentation
        self.min_len, self.max_len = min_len, max_len
        self.train_set = self.val_set = self.test_set = None

    def setup(self, stage: Union[str, None] = None) -> None:
        if stage in ("fit", None):
            assert isinstance(self.split_ratio, float) and self.split_ratio > 0, \
                "Invalid split ratio. Must be a float between 0 and 1."
            raw = RNADataset.from_path(self.data_path)
            if self.noise_augmentation is not None:
                raw.noise_augmentation(num_gen=self.noise_augmentation)
            if self.slice_augmentation is not None:
                raw.slice_augmentation(num_gen=self.slice_augmentation)
            if self.min_len is not None:
                raw.filter_by_min_length(min_len=self.min_len)
            if self.max_len is not None:
                raw.filter_by_max_length(max_len=self.max_len)
            self.train_set, self.val_set = self._split_dataset(raw, split_ratio=self.split_ratio)
        if stage in ("test", None):
            self.test_set = RNADataset.from_path(self.data_path, is_predict=True)

    def _loader(self, dataset, shuffle: bool) -> DataLoader:
        return DataLoader(dataset, batch_size=self.batch_size, num_workers=self.num_workers, shuffle=shuffle,
                          persistent_workers=self.num_workers > 0, collate_fn=self._featurize)

    def train_dataloader(self) -> DataLoader:
        return self._loader(self.train_set, True)

    def val_dataloader(self) -> DataLoader:
        return self._loader(self.val_set, False)

    def test_dataloader(self) -> DataLoader:
        return self._loader(self.test_set, False)

    _featurize = staticmethod(featurize)

    @staticmethod
    def _split_dataset(dataset: RNADataset, split_ratio: Union[float, List[float]]) -> Tuple[RNADataset, ...]:
        if isinstance(split_ratio, float):
            if not 0 < split_ratio < 1:
                raise ValueError("Proportion must be a float between 0 and 1.")
            fractions = [split_ratio, 1 - split_ratio]
        elif isinstance(split_ratio, list):
            if not all(isinstance(p, float) and p > 0 for p in split_ratio):
                raise ValueError("All split_ratio must be positive floats.")
            if abs(sum(split_ratio) - 1) >= 1e-6:
                raise ValueError("The sum of split_ratio must equal 1.")
            fractions = split_ratio
        else:
            raise TypeError("split_ratio must be a float or a list of floats.")
        by_id: Dict[str, List[Item]] = {}
        for item in dataset.data:
            by_id.setdefault(item['id'], []).append(item)
        ids = list(by_id)
        random.shuffle(ids)
        sizes = [int(len(ids) * p) for p in fractions]
        sizes[-1] += len(ids) - sum(sizes)
        parts, at = [], 0
        for n in sizes:
            parts.append(RNADataset([item for rna_id in ids[at:at + n] for item in by_id[rna_id]]))
            at += n
        return tuple(parts)


def gen_seq_csv(data_path: str = f'{DATA_PATH}seqs/', output_path: str = f'{DATA_PATH}ref.csv'):
    """pdb_id,seq table of every FASTA record under ``data_path`` (data.py:541-557)."""
    import pandas as pd
    rows = [rec for fname in os.listdir(data_path) for rec in read_fasta(os.path.join(data_path, fname))]
    pd.DataFrame({"pdb_id": [r[0] for r in rows], "seq": [r[1] for r in rows]}).to_csv(output_path, index=False)
    print(f"CSV file saved at {output_path}")


def cal_recovery_rate(pred_path: str, ref_path: str, output_path: str = f'{DATA_PATH}recovery.csv'):
    """Per-structure recovery (matching positions / reference length) of a prediction CSV (data.py:559-579)."""
    import pandas as pd
    merged = pd.merge(pd.read_csv(ref_path), pd.read_csv(pred_path), on='pdb_id', suffixes=('_ref', '_pred'))
    rows = [{'pdb_id': r['pdb_id'], 'length': len(r['seq_ref']),
             'recovery_rate': sum(a == b for a, b in zip(r['seq_ref'], r['seq_pred'])) / len(r['seq_ref'])}
            for _, r in merged.iterrows()]
    pd.DataFrame(rows, columns=['pdb_id', 'recovery_rate', 'length']).to_csv(output_path, index=False)
    print(f"Recovery rates saved at {output_path}")


def separate(concat: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """[sum L] -> zero-padded [B, max L] by the per-structure lengths."""
    lens = [int(v) for v in lengths.tolist()]
    out = torch.zeros(len(lens), max(lens) if lens else 0, dtype=concat.dtype, device=concat.device)
    at = 0
    for b, L in enumerate(lens):
        out[b, :L] = concat[at:at + L]
        at += L
    return out
