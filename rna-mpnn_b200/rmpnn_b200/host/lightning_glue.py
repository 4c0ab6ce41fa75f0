"""Trainer factory and callbacks behind ``{rdesign,rnampnn}.utils.train`` (reference: rdesign/utils/train.py:11-115,
rnampnn/utils/train.py:13-117).  The callbacks only read the per-batch records the LightningModules accumulate in
``val_step_outputs`` / ``test_step_outputs``; pytorch_lightning is imported lazily and only ``get_trainer`` needs it."""
from __future__ import annotations

import pickle

import numpy as np
import torch

try:
    import pytorch_lightning as pl
    Callback = pl.Callback
except ImportError:                                   # this image has no Lightning: the classes stay importable
    pl = None

    class Callback:                                   # type: ignore[no-redef]
        pass


def _ratio(records, num: str, den: str, device):
    total = torch.as_tensor(records[den], dtype=torch.float32).sum()
    return (torch.as_tensor([float(v) for v in records[num]], dtype=torch.float32).sum() / total).to(device)


class LossMonitor(Callback):
    """Epoch aggregates of loss and recovery: token-weighted and per-structure mean (train.py:21-49)."""

    @staticmethod
    def _flush(model, attr: str, prefix: str):
        rec = getattr(model, attr)
        model.log(f'{prefix}_loss', _ratio(rec, f'{prefix}_loss', 'len', model.device), prog_bar=True, sync_dist=True)
        model.log(f'weighted_{prefix}_recovery_rate', _ratio(rec, 'correct', 'len', model.device), prog_bar=True,
                  sync_dist=True)
        per_structure = torch.as_tensor([float(v) for v in rec['recovery_rates']], dtype=torch.float32)
        model.log(f'{prefix}_recovery_rate', per_structure.mean().to(model.device), prog_bar=True, sync_dist=True)
        setattr(model, attr, {f'{prefix}_loss': [], 'correct': [], 'len': [], 'recovery_rates': []})

    def on_validation_epoch_end(self, trainer, model):
        self._flush(model, 'val_step_outputs', 'val')

    def on_test_epoch_end(self, trainer, model):
        self._flush(model, 'test_step_outputs', 'test')


class NameModel(Callback):
    """Stamps run name / version on the module at fit start (they go into the checkpoint, rdesign.py:174-182)."""

    def __init__(self, name: str, version: int):
        super().__init__()
        self.name, self.version = name, version

    def on_fit_start(self, trainer, model) -> None:
        model.name, model.version = self.name, self.version


class XGBTrainerBase(Callback):
    """After ``fit``: reload the best checkpoint, embed the training set, fit the XGBoost head on the embeddings,
    score it on the validation loader and pickle it next to the checkpoint (train.py:52-89).  Subclasses say how a
    batch becomes (embedding rows, class ids)."""
    output_path = './out/'

    def model_class(self):
        raise NotImplementedError

    def embed_batch(self, model, batch):
        raise NotImplementedError

    def collect(self, loader, model):
        feats, labels = [], []
        for batch in loader:
            with torch.no_grad():
                x, y = self.embed_batch(model, batch)
            feats.append(x.detach().float().cpu().numpy())
            labels.append(y.detach().cpu().numpy().astype(np.float64))
        if not feats:
            return np.zeros((0, 0)), np.zeros(0)
        return np.concatenate(feats, 0), np.concatenate(labels, 0)

    # the reference names this helper _generate_embedding(dataloader, model)
    def _generate_embedding(self, dataloader, model):
        return self.collect(dataloader, model)

    def on_fit_end(self, trainer, model) -> None:
        ckpt_dir = f"{self.output_path}checkpoints/{model.name}/"
        best = self.model_class().load_from_checkpoint(f"{ckpt_dir}Final-V{model.version}.ckpt")
        print('=' * 20, '\nStart training XGBoost!\n')
        X, Y = self.collect(trainer.train_dataloader, best)
        best.xgb_readout.fit(X, Y)
        print(f'Training score: {best.xgb_readout.score(X, Y)}\n')
        Xv, Yv = self.collect(trainer.val_dataloaders, best)
        print(f'Validation score: {best.xgb_readout.score(Xv, Yv)}\n')
        print('XGBoost training done!\n', '=' * 20)
        with open(f"{ckpt_dir}XGB-V{best.version}.pkl", 'wb') as f:
            pickle.dump(best.xgb_readout, f)


def build_trainer(output_path: str, name: str, version: int, max_epochs: int, val_check_interval: int, callbacks,
                  logs_dir: str, progress_bar: bool = True, verbose_ckpt: bool = False, **trainer_kwargs):
    """``pl.Trainer`` with the reference's logger / checkpoint policy: TensorBoard under ``logs_dir``, the best
    ``val_recovery_rate`` checkpoint as ``Final-V<version>.ckpt``, one process per visible GPU."""
    if pl is None:
        raise ImportError("get_trainer needs pytorch_lightning (training orchestration is host code outside the "
                          "B200 hot path and is not vendored)")
    from pytorch_lightning.callbacks import ModelCheckpoint
    from pytorch_lightning.loggers import TensorBoardLogger
    logger = TensorBoardLogger(save_dir=logs_dir, name=name, version=version)
    checkpoint = ModelCheckpoint(dirpath=f"{output_path}checkpoints/{name}/", filename=f'Final-V{version}',
                                 save_top_k=1, verbose=verbose_ckpt, monitor='val_recovery_rate', mode='max')
    return pl.Trainer(accelerator="auto", devices=-1 if torch.cuda.is_available() else 1, max_epochs=max_epochs,
                      enable_progress_bar=progress_bar, logger=logger, callbacks=[checkpoint] + list(callbacks),
                      check_val_every_n_epoch=val_check_interval, **trainer_kwargs)
