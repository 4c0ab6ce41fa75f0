"""Host-side data plumbing of the rdesign package (out of the hot path, plain
numpy/torch).  Mirrors the public names of rdesign/utils/data.py: the collate
``featurize`` defines the padded-batch layout the kernels consume (:84-119), and
``normalize`` / ``separate`` / ``concat`` / ``gen_mask`` are helpers other modules import.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch.utils.data import Dataset

from ..config.glob import DATA_PATH, SPLIT_RATIO

ATOMS = ["P", "O5'", "C5'", "C4'", "C3'", "O3'"]
ALPHABET = 'AUCG'

try:                                               # optional third-party pieces
    from pytorch_lightning import LightningDataModule
except ImportError:                                # pragma: no cover - image without Lightning
    class LightningDataModule:                     # minimal stand-in so the class below still imports
        pass


def _read_fasta(path):
    name, seq = None, []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith('>'):
                if name is not None:
                    break
                name = line[1:].split()[0]
            elif line:
                seq.append(line)
    return {} if name is None else {name: ''.join(seq)}


def read_fasta_biopython(file_path):
    return _read_fasta(file_path)


def gen_dataframe(file_path=DATA_PATH + 'seqs/'):
    import pandas as pd
    rows = []
    for fn in os.listdir(file_path):
        rec = _read_fasta(os.path.join(file_path, fn))
        if not rec:
            print(f"Error reading file {fn}: {rec}")
            continue
        rows.extend(rec.items())
    return pd.DataFrame({"pdb_id": [r[0] for r in rows], "seq": [r[1] for r in rows]})


def split_dataset(data, output_dir=DATA_PATH, split_ratio=SPLIT_RATIO):
    data['split'] = np.random.choice(['train', 'valid', 'test'], size=len(data), p=split_ratio)
    for part in ('train', 'valid', 'test'):
        data[data['split'] == part].to_csv(f"{output_dir}{part}_data.csv", index=False)


class RNADataset(Dataset):
    """csv (pdb_id, seq) + <npy_dir>/<pdb_id>.npy [L,7,3] -> feature dict per structure."""

    def __init__(self, data_path, npy_dir):
        super().__init__()
        import pandas as pd
        self.data = pd.read_csv(data_path)
        self.npy_dir = npy_dir
        self.seq_list = self.data['seq'].to_list()
        self.name_list = self.data['pdb_id'].to_list()

    def __len__(self):
        return len(self.name_list)

    def __getitem__(self, idx):
        coords = np.load(os.path.join(self.npy_dir, self.name_list[idx] + '.npy'))
        return {"name": self.name_list[idx], "seq": self.seq_list[idx],
                "coords": {a: coords[:, i, :] for i, a in enumerate(ATOMS)}}


def featurize(batch):
    """Collate: X [B,N,6,3] f32 (NaN atoms -> 0, zero padded), S [B,N] i64, mask [B,N] f32
    prefix mask, lengths int32 [B], names."""
    B = len(batch)
    lengths = np.array([len(b['seq']) for b in batch], dtype=np.int32)
    N = int(lengths.max())
    X = np.zeros((B, N, 6, 3), np.float32)
    S = np.zeros((B, N), np.int64)
    mask = np.zeros((B, N), np.float32)
    names = []
    for i, b in enumerate(batch):
        x = np.nan_to_num(np.stack([b['coords'][a] for a in ATOMS], 1), nan=0.0)
        keep = np.isfinite(x.sum((1, 2)))            # residues with infinite coordinates are dropped
        n = int(keep.sum())
        X[i, :n] = x[keep]
        S[i, :n] = np.asarray([ALPHABET.index(c) for c in b['seq']], np.int64)[keep]
        mask[i, :n] = 1.0
        names.append(b['name'])
    return torch.from_numpy(X), torch.from_numpy(S), torch.from_numpy(mask), lengths, names


class RNADataModule(LightningDataModule):
    num_workers = 19                                   # the reference's loader setting; tests set 0

    @classmethod
    def from_defaults(cls, batch_size=16, split_ratio=SPLIT_RATIO):
        split_dataset(gen_dataframe(), split_ratio=split_ratio)
        return cls(*(DATA_PATH + f'{p}_data.csv' for p in ('train', 'valid', 'test')),
                   './data/coords', './data/coords', './data/coords', batch_size=batch_size)

    def __init__(self, train_data_path, valid_data_path, test_data_path, train_npy_dir, valid_npy_dir,
                 test_npy_dir, batch_size=32):
        super().__init__()
        self.paths = {'train': (train_data_path, train_npy_dir), 'valid': (valid_data_path, valid_npy_dir),
                      'test': (test_data_path, test_npy_dir)}
        self.batch_size = batch_size

    def setup(self, stage=None):
        if stage in ('fit', None):
            self.train_dataset = RNADataset(*self.paths['train'])
            self.valid_dataset = RNADataset(*self.paths['valid'])
        if stage in ('test', None):
            self.test_dataset = RNADataset(*self.paths['test'])

    def _loader(self, ds, shuffle):
        return torch.utils.data.DataLoader(ds, batch_size=self.batch_size, num_workers=self.num_workers, shuffle=shuffle,
                                           persistent_workers=self.num_workers > 0, collate_fn=featurize)

    def train_dataloader(self):
        return self._loader(self.train_dataset, True)

    def val_dataloader(self):
        return self._loader(self.valid_dataset, False)

    def test_dataloader(self):
        return self._loader(self.test_dataset, False)


def nan_to_num(tensor, nan=0.0):
    tensor[torch.isnan(tensor)] = nan
    return tensor


def normalize(tensor, dim=-1):
    """v / |v| with NaN (0/0) -> 0: the convention the featurizer kernels copy."""
    return nan_to_num(torch.div(tensor, torch.norm(tensor, dim=dim, keepdim=True)))


def _lengths_list(lengths):
    return [int(v) for v in (lengths.tolist() if hasattr(lengths, 'tolist') else lengths)]


def separate(concat: torch.Tensor, lengths) -> torch.Tensor:
    """[sum L, ...] -> zero padded [B, Lmax, ...] (CPU tensor, like the reference)."""
    ls = _lengths_list(lengths)
    out = torch.zeros((len(ls), max(ls)) + tuple(concat.shape[1:]), dtype=concat.dtype)
    for i, chunk in enumerate(torch.split(concat.detach().cpu(), ls)):
        out[i, :ls[i]] = chunk
    return out


def concat(separated: torch.Tensor, lengths) -> torch.Tensor:
    ls = _lengths_list(lengths)
    return torch.cat([separated[i, :n] for i, n in enumerate(ls)], 0)


def gen_mask(lengths) -> torch.Tensor:
    ls = torch.as_tensor(_lengths_list(lengths))
    return (torch.arange(int(ls.max()))[None, :] < ls[:, None]).float()


def cal_recovery_rate(pred_path: str, ref_path: str, output_path: str = f'{DATA_PATH}recovery.csv'):
    import pandas as pd
    m = pd.merge(pd.read_csv(ref_path), pd.read_csv(pred_path), on='pdb_id', suffixes=('_ref', '_pred'))
    rec = [{'pdb_id': r.pdb_id, 'length': len(r.seq_ref),
            'recovery_rate': sum(a == b for a, b in zip(r.seq_ref, r.seq_pred)) / len(r.seq_ref)}
           for r in m.itertuples()]
    pd.DataFrame(rec).to_csv(output_path, index=False)
    print(f"Recovery rates saved at {output_path}")
