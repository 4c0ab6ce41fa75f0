"""Drop-in ``rdesign.utils.train`` (reference: rdesign/utils/train.py): ``get_trainer`` with the reference's exact
signature, and the callbacks ``LossMonitor`` / ``NameModel`` / ``XGBTrainer`` under their reference names.  Host
orchestration: needs pytorch_lightning (and xgboost for ``train_xgb=True``) at call time, not at import time."""
from rmpnn_b200.host.lightning_glue import LossMonitor, NameModel, XGBTrainerBase, build_trainer
from ..config.glob import OUTPUT_PATH


class XGBTrainer(XGBTrainerBase):
    """Fits ``RNAModel.xgb_readout`` on ``h_V`` of the whole training set (train.py:52-89)."""
    output_path = OUTPUT_PATH

    def model_class(self):
        from ..model.rdesign import RNAModel
        return RNAModel

    def embed_batch(self, model, batch):
        X, S, mask, _lengths, _names = batch
        h_V, S_valid = model(X.to(model.device), S.to(model.device), mask.to(model.device), is_predict=True)
        return h_V, S_valid


def get_trainer(name: str, version: int, max_epochs: int = 30, val_check_interval: int = 1, train_xgb: bool = False):
    callbacks = [LossMonitor(), NameModel(name, version)] + ([XGBTrainer()] if train_xgb else [])
    return build_trainer(OUTPUT_PATH, name, version, max_epochs, val_check_interval, callbacks,
                         logs_dir=f"{OUTPUT_PATH}logs", verbose_ckpt=True)
