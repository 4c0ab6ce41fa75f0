"""Drop-in ``rnampnn.utils.predict.predict`` (reference: rnampnn/utils/predict.py:11-29): competition directory ->
``model.predict`` per batch -> ``submit.csv``.  Host I/O around the kernels."""
import os

import torch

from ..config.glob import COMPETITION_DATA, COMPETITION_OUT
from .data import RNADataset, featurize


def predict(model, batch_size=32, data_path=COMPETITION_DATA, output_path=COMPETITION_OUT, filename="submit.csv",
            num_workers=19) -> None:
    import pandas as pd
    dataset = RNADataset.from_path(data_path=data_path, is_predict=True)
    loader = torch.utils.data.DataLoader(dataset, batch_size, num_workers=num_workers, shuffle=False,
                                         persistent_workers=num_workers > 0, collate_fn=featurize)
    scratch = f'./{filename}'
    for batch_id, batch in enumerate(loader):
        model.predict(batch, batch_id, output_dir='./', filename=filename)
    pd.read_csv(scratch).to_csv(f"{output_path}{filename}", index=False)
    os.remove(scratch)
    print(f"Predictions saved to {output_path}{filename}")
