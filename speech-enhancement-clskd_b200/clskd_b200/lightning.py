"""`KnowledgeDistillation`: the reference's training module (distill.py:38-229 and its SPKD / MSE / STFT / ReviewKD
siblings) over the B200-native step, with the same hook names - training_step, validation_step,
configure_optimizers - so it drops into a PyTorch-Lightning Trainer when Lightning is installed, and into the plain
`fit` loop below when it is not (this image has no Lightning; nothing here imports it unless present).

Differences from the reference module that are deliberate:
  * training_step runs the fused step of `distill.DistillTrainer` (flat-bucket gradient pack, NCCL all-reduce
    across ranks, fused Adam) under MANUAL optimisation (`automatic_optimization = False`); the reference lets
    Lightning call backward() and torch.optim.Adam;
  * validation_step enhances the WHOLE batch at once on the GPU (streaming inference for long utterances) and
    computes batched SI-SDR / SNR metrics there (`metrics.batch_metrics`); the reference loops over utterances and
    calls asteroid's CPU `get_metrics` for each (distill.py:159-187) - STOI / PESQ stay CPU libraries outside the path;
  * the teacher runs in eval mode under no_grad unless `faithful=True`.
"""
import torch
import torch.nn as nn

from . import config as cfg
from .distill import DistillTrainer
from .metrics import batch_metrics

try:                                        # optional: only when the host framework is installed
    import pytorch_lightning as _pl
    _Base = _pl.LightningModule
except Exception:                           # noqa: BLE001
    try:
        import lightning.pytorch as _pl
        _Base = _pl.LightningModule
    except Exception:                       # noqa: BLE001
        _pl = None
        _Base = nn.Module


class KnowledgeDistillation(_Base):
    def __init__(self, teacher, student, mode="clskd", lr=cfg.learning_rate, weight_decay=0.0, faithful=False,
                 fresh_abf=False, base="stft", stream_chunk_frames=400):
        super().__init__()
        self.teacher, self.student = teacher, student
        for p in self.teacher.parameters():                       # distill.py:49-50
            p.requires_grad = False
        self.mode, self.lr, self.weight_decay = mode, lr, weight_decay
        self.faithful, self.fresh_abf, self.base = faithful, fresh_abf, base
        self.stream_chunk_frames = stream_chunk_frames
        self.automatic_optimization = False                        # the fused step owns backward + Adam
        self._trainer_impl = None
        self.logged = {}                                           # last logged values when no Lightning logger exists

    # ---- helpers
    def _impl(self, X):
        if self._trainer_impl is None:
            self._trainer_impl = DistillTrainer(self.teacher, self.student, mode=self.mode, lr=self.lr,
                                                weight_decay=self.weight_decay, faithful=self.faithful,
                                                fresh_abf=self.fresh_abf, base=self.base, example_input=X[:2])
        return self._trainer_impl

    def _log(self, name, value, **kw):
        self.logged[name] = float(value)
        if _pl is not None and getattr(self, "_trainer", None) is not None:
            self.log(name, value, **kw)

    @staticmethod
    def _wave(t):
        return t.squeeze(1) if t.dim() == 3 else t

    # ---- Lightning hooks (same names / batch format as the reference: batch = (mixture X, clean y))
    def forward(self, x):
        return self.student(self._wave(x), is_feat=True)

    def training_step(self, batch, batch_idx=0):
        X, y = (self._wave(t) for t in batch)
        loss = self._impl(X).train_step(X, y)
        # device scalars: logging must not synchronise the hot path (read them with float() when needed)
        self.logged["train_loss"] = loss
        for k, v in self._impl(X).step_fn.last_terms.items():
            self.logged["train_" + k] = v.detach()
        if _pl is not None and getattr(self, "_trainer", None) is not None:
            self.log("train_loss", loss, prog_bar=True)
        return loss

    def validation_step(self, batch, batch_idx=0):
        X, y = (self._wave(t) for t in batch)
        was_training = self.student.training
        self.student.eval()
        with torch.no_grad():
            T = X.shape[-1] // self.student.win_inc + 3
            if T > 4 * self.stream_chunk_frames and hasattr(self.student, "enhance_streaming"):
                est = self.student.enhance_streaming(X, self.stream_chunk_frames)
            else:
                est = self.student(X, is_feat=True)
            L = est.shape[-1]
            means, per = batch_metrics(X[:, :L], y[:, :L], est)
        self.student.train(was_training)
        for k, v in means.items():
            self._log(k, v, on_epoch=True)
        return means

    def configure_optimizers(self):
        """The fused flat-bucket Adam lives inside the step (manual optimisation).  For a host framework that insists
        on an optimizer object, the reference's own choice (distill.py:202-204) over the same parameters is returned;
        it is never stepped by this module."""
        params = [p for p in self.student.parameters() if p.requires_grad]
        return torch.optim.Adam(params, lr=self.lr, weight_decay=self.weight_decay)

    # ---- plain loop for hosts without Lightning
    def fit(self, train_batches, val_batches=None, epochs=1):
        """minimal Trainer.fit: `train_batches` / `val_batches` are iterables of (X, y) device batches"""
        history = []
        for ep in range(epochs):
            last = None
            for i, batch in enumerate(train_batches):
                last = self.training_step(batch, i)
            rec = {"epoch": ep, "train_loss": float(last) if last is not None else None}
            if val_batches is not None:
                agg, n = {}, 0
                for i, batch in enumerate(val_batches):
                    m = self.validation_step(batch, i)
                    n += 1
                    for k, v in m.items():
                        agg[k] = agg.get(k, 0.0) + v
                rec.update({k: v / max(n, 1) for k, v in agg.items()})
            history.append(rec)
        return history
