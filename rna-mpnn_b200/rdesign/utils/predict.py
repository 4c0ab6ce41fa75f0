"""``predict`` with the reference's signature (rdesign/utils/predict.py:10-30): loads the
competition directory, runs ``model.predict`` per batch, writes ``submit.csv``.  Host I/O."""
import os

import torch

from ..config.glob import COMPETITION_DATA, COMPETITION_OUT
from ..utils.data import RNADataset, featurize, gen_dataframe


def predict(model, batch_size=32, data_path=COMPETITION_DATA, output_path=COMPETITION_OUT, filename="submit.csv",
            num_workers=19) -> None:
    import pandas as pd
    gen_dataframe(file_path=f'{data_path}seqs/').to_csv('./predict_data.csv', index=False)
    dataset = RNADataset('./predict_data.csv', f'{data_path}coords')
    loader = torch.utils.data.DataLoader(dataset, batch_size, num_workers=num_workers, shuffle=False,
                                         persistent_workers=num_workers > 0, collate_fn=featurize)
    if os.path.exists(f'./{filename}'):
        os.remove(f'./{filename}')
    for batch_id, batch in enumerate(loader):
        model.predict(batch, batch_id, output_dir='./', filename=filename)
    pd.read_csv(f"./{filename}").to_csv(f"{output_path}{filename}", index=False)
    os.remove('./predict_data.csv')
    os.remove(f'./{filename}')
    print(f"Predictions saved to {output_path}{filename}")
