"""Drop-in ``rnampnn.utils.train`` (reference: rnampnn/utils/train.py): ``get_trainer`` (bf16-mixed, DDP with unused
parameters allowed, XGBoost head fitted after ``fit``) and the callbacks under their reference names.  Host
orchestration: needs pytorch_lightning / xgboost at call time only."""
import torch

from rmpnn_b200.host.lightning_glue import LossMonitor, NameModel, XGBTrainerBase, build_trainer
from ..config.glob import OUTPUT_PATH


class XGBTrainer(XGBTrainerBase):
    """Fits ``RNAMPNN.xgb_readout`` on ``embedding(coords, mask)`` of the valid residues (train.py:53-89)."""
    output_path = OUTPUT_PATH

    def model_class(self):
        from ..model.rnampnn import RNAMPNN
        return RNAMPNN

    def embed_batch(self, model, batch):
        sequences, coords, mask, _ids = batch
        sequences, coords, mask = (t.to(model.device) for t in (sequences, coords, mask))
        keep = mask.bool()
        return model.embedding(coords, mask)[keep], torch.argmax(sequences[keep], dim=-1)


def get_trainer(name: str, version: int, max_epochs: int = 60, val_check_interval: int = 1, progress_bar=True):
    return build_trainer(OUTPUT_PATH, name, version, max_epochs, val_check_interval,
                         [LossMonitor(), XGBTrainer(), NameModel(name, version)], logs_dir=f"{OUTPUT_PATH}/logs",
                         progress_bar=progress_bar, precision="bf16-mixed",
                         strategy='ddp_find_unused_parameters_true', enable_checkpointing=True)
