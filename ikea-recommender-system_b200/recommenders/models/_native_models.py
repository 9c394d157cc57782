"""Train-step drivers shared by the drop-in trainers: stage the batch, call the C ABI."""

from __future__ import annotations

import random

import torch

from .._base import make_hparams


def supervised_train_step(trainer, s, a, true_len) -> torch.Tensor:
    """GRU4Rec_trainer / BidirGRU4Rec_trainer.train_step (GRU4Rec/model.py:129-155)."""
    net = trainer.gru_model
    if net._family == "bidir" and net.training and net.dropout.p > 0:
        raise NotImplementedError("BidirGRU4Rec dropout > 0 in training mode is not implemented natively yet")
    B = int(s.shape[0])
    eng = trainer._ready(B)
    ds, _, da, dln, _, _, _ = trainer._stager.stage(s, a, true_len)
    batch = eng._batch(B, ds, da, dln)
    hp = make_hparams(trainer.learning_rate)
    eng.train_step_supervised(batch, hp, trainer._loss_dev)
    return trainer._loss_dev[0]


def pick_main(trainer):
    """The double-Q coin flip, consuming the python RNG exactly like sqn_gru.py:207-216."""
    return 0 if random.uniform(0, 1) <= 0.5 else 1


def q_train_step(trainer, hp, s, a, r, s_next, true_len, true_next_len, is_end, main=None) -> torch.Tensor:
    """SQN_trainer / SMORL_trainer.train_step; returns device tensor [sup_loss, q_loss]."""
    B = int(s.shape[0])
    eng = trainer._ready(B)
    if main is None:
        main = pick_main(trainer)
    trainer.last_main = main + 1
    ds, dsn, da, dln, dnl, dr, de = trainer._stager.stage(s, a, true_len, r, s_next, true_next_len, is_end)
    batch = eng._batch(B, ds, da, dln, dr, dsn, dnl, de)
    eng.train_step_q(batch, hp, main, trainer._loss_dev)
    return trainer._loss_dev[:2]
