"""``get_trainer`` with the reference's signature (rdesign/utils/train.py:91-115).

Training orchestration (Lightning Trainer, TensorBoard logger, checkpoint callback on
``val_recovery_rate``) is outside the B200 hot path and is delegated to
pytorch_lightning, which must be installed to use it."""
from ..config.glob import OUTPUT_PATH


def get_trainer(name: str, version: int, max_epochs: int = 100, train_xgb: bool = False, **trainer_kwargs):
    try:
        import pytorch_lightning as pl
        from pytorch_lightning.callbacks import ModelCheckpoint
        from pytorch_lightning.loggers import TensorBoardLogger
    except ImportError as e:
        raise ImportError("rdesign.utils.train.get_trainer needs pytorch_lightning (training orchestration is "
                          "not part of the B200 hot path)") from e

    class LossMonitor(pl.Callback):
        """Aggregates the per-batch validation records RNAModel accumulates."""

        def on_validation_epoch_end(self, trainer, module):
            out = module.val_step_outputs
            if out['len']:
                total = sum(out['len'])
                module.log('val_loss', sum(float(v) for v in out['val_loss']) / total, sync_dist=True)
                module.log('weighted_val_recovery_rate', sum(out['correct']) / total, sync_dist=True)
                module.log('val_recovery_rate', sum(out['recovery_rates']) / len(out['recovery_rates']),
                           sync_dist=True)
            for v in out.values():
                v.clear()

    class NameModel(pl.Callback):
        def on_fit_start(self, trainer, module):
            module.name, module.version = name, version

    logger = TensorBoardLogger(save_dir=f'{OUTPUT_PATH}logs', name=name, version=version)
    ckpt = ModelCheckpoint(dirpath=f'{OUTPUT_PATH}checkpoints/{name}/', filename=f'Final-V{version}', save_top_k=1,
                           monitor='val_recovery_rate', mode='max')
    return pl.Trainer(max_epochs=max_epochs, logger=logger, callbacks=[ckpt, LossMonitor(), NameModel()],
                      accelerator='gpu', devices=-1, **trainer_kwargs)
