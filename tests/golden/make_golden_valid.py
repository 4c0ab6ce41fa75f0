"""Generate tests/golden/valid_set.npz: the reference's own held-out structures (data/valid_data.csv; data/test_data.csv
is empty) run one by one (batch 1, the reference's inference setting) through the UNMODIFIED
reference RNAModel on CPU with the seeded default weights (all 234 structures).  Stores the packed coordinates (NaNs kept: NaN -> 0 is part of
the collate under test), the true sequences, and the reference's logits / argmax sequences.

    python tests/golden/make_golden_valid.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import REF, load_reference_model  # noqa: E402
from oracle.weights import RDesignConfig  # noqa: E402


@torch.no_grad()
def main():
    import pandas as pd
    torch.set_num_threads(8)
    df = pd.read_csv(os.path.join(REF, "data", "valid_data.csv"))
    keep = list(df.index)                                 # all 234 held-out structures (123,781 nt, L 2..3065)
    model = load_reference_model(RDesignConfig(), 42)
    coords, lengths, seqs, logits, names = [], [], [], [], []
    for i in keep:
        name, seq = df.loc[i, "pdb_id"], df.loc[i, "seq"]
        c = np.load(os.path.join(REF, "data", "coords", name + ".npy"))[:, :6, :].astype(np.float32)
        # reference collate (rdesign/utils/data.py:84-119): NaN -> 0, residues with infinite coordinates dropped
        x = np.nan_to_num(c, nan=0.0)
        assert np.isfinite(x).all() and len(seq) == len(c)
        X = torch.from_numpy(x)[None]
        S = torch.tensor(["AUCG".index(ch) for ch in seq])[None]
        mask = torch.ones(1, len(c))
        h_V, _ = model(X, S, mask)
        lg = model.readout(h_V)
        coords.append(c); lengths.append(len(c)); seqs.append(S[0].numpy().astype(np.int8))
        logits.append(lg.numpy()); names.append(name)
    lg = np.concatenate(logits, 0)
    out = dict(coords=np.concatenate(coords, 0), lengths=np.array(lengths, np.int32),
               seq=np.concatenate(seqs, 0), logits=lg, samples=lg.argmax(-1).astype(np.int8),
               names=np.array(names), seed=np.int64(42))
    path = os.path.join(HERE, "valid_set.npz")
    np.savez_compressed(path, **out)
    rec = float((out["samples"] == out["seq"]).mean())
    print(f"{len(keep)} structures, {int(sum(lengths))} nt, recovery of the seeded random weights {rec:.4f} "
          f"-> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
