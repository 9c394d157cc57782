"""Train-step drivers shared by the drop-in trainers: stage the batch, call the C ABI."""

from __future__ import annotations

import random

import torch

from .._base import make_hparams


def _supervised_hparams(trainer):
    """Adam hyper-parameters + the BidirGRU4Rec dropout of the supervised step (train mode only)."""
    net = trainer.gru_model
    hp = make_hparams(trainer.learning_rate)
    if net._family == "bidir" and net.training and net.dropout.p > 0:
        hp.dropout_p = float(net.dropout.p)
        hp.dropout_seed = int(getattr(trainer, "_dropout_seed", 118))
        # device uint8 [B, 2H], 1 = keep (parity tests); vocabulary-sharded trainer: the mask of the GLOBAL batch
        # [world * B, 2H], rows in rank order (every rank drops the same elements, no collective needed)
        mask = getattr(trainer, "dropout_mask_override", None)
        if mask is not None:
            trainer._dropout_mask_keepalive = mask = mask.to(device=net._param_device(), dtype=torch.uint8).contiguous()
            hp.dropout_mask = mask.data_ptr()
    return hp


def supervised_train_step(trainer, s, a, true_len) -> torch.Tensor:
    """GRU4Rec_trainer / BidirGRU4Rec_trainer.train_step (GRU4Rec/model.py:129-155)."""
    B = int(s.shape[0])
    hp = _supervised_hparams(trainer)
    if trainer._shard is not None:
        return _sharded_step(trainer, hp, 0, s, a, true_len)[0]
    eng = trainer._ready(B)
    ds, _, da, dln, _, _, _ = trainer._stager.stage(s, a, true_len)
    batch = eng._batch(B, ds, da, dln)
    eng.train_step_supervised(batch, hp, trainer._loss_dev)
    return trainer._loss_dev[0]


def _host(t, dtype):
    """CPU tensor in the dtype/layout the C ABI reads (no copy in the common case)."""
    if t.dtype != dtype:
        t = t.to(dtype)
    return t if t.is_contiguous() else t.contiguous()


def host_path(trainer, s) -> bool:
    """CPU tensors on an unsharded trainer go through the synchronous host entry points of the C ABI."""
    return (not s.is_cuda) and trainer._shard is None


def supervised_train_step_host(trainer, s, a, true_len) -> float:
    """GRU4Rec_trainer / BidirGRU4Rec_trainer.train_step with CPU tensors: rec_train_step_supervised_host."""
    B = int(s.shape[0])
    eng = trainer._ready(B)
    hs, ha, hl = _host(s, torch.int64), _host(a, torch.int64), _host(true_len, torch.int64)
    trainer._stager.h2d_bytes = eng.host_batch_bytes
    return eng.train_step_supervised_host(eng._batch(B, hs, ha, hl), _supervised_hparams(trainer))


def q_train_step_host(trainer, hp, s, a, r, s_next, true_len, true_next_len, is_end, main=None):
    """SQN_trainer / SMORL_trainer.train_step with CPU tensors: rec_train_step_q_host -> (sup_loss, q_loss)."""
    B = int(s.shape[0])
    if main is None:
        main = pick_main(trainer)
    trainer.last_main = main + 1
    eng = trainer._ready(B)
    hs, ha, hl = _host(s, torch.int64), _host(a, torch.int64), _host(true_len, torch.int64)
    hsn, hnl = _host(s_next, torch.int64), _host(true_next_len, torch.int64)
    hr = _host(r, torch.float32).reshape(-1)  # (q6) rewards are cast to fp32
    he = _host(is_end, torch.uint8)
    trainer._stager.h2d_bytes = eng.host_batch_bytes
    return eng.train_step_q_host(eng._batch(B, hs, ha, hl, hr, hsn, hnl, he), hp, main)


def _sharded_step(trainer, hp, main, s, a, true_len, r=None, s_next=None, true_next_len=None, is_end=None):
    """Vocabulary-sharded step: ONE all-gather of the packed local batches, then phases A-D with their
    three collectives (records all-gather, Q all-reduce, dh all-reduce)."""
    import ctypes as C
    import torch.distributed as dist
    from ...sharded import ShardedStep
    from ... import _native as N
    rank, world, group = trainer._shard
    B = int(s.shape[0])
    Bg = B * world
    eng = trainer._ready(Bg)
    ds, dsn, da, dln, dnl, dr, de = trainer._stager.stage(s, a, true_len, r, s_next, true_next_len, is_end)
    L = int(ds.shape[1])
    st = getattr(trainer, "_sharded_step", None)
    if st is None or st.eng is not eng or st.B_local != B:
        n0 = trainer._nets[0]
        st = trainer._sharded_step = ShardedStep(eng, world, group, n0.hidden_dim * (2 if n0._bidirectional else 1), rank=rank,
                                                 shard_embedding=n0._emb_shard is not None)
        st.alloc_inputs(B, L, ds.device)
    local = eng._batch(B, ds, da, dln, dr, dsn, dnl, de)
    st.step(local, hp, main, trainer._loss_dev, has_q=r is not None)
    if st.shard_embedding:  # the owners moved every row of the trained table: the other rows of this rank's copy are stale
        trainer._nets[main]._emb_stale = trainer._nets[main]._emb_moments_stale = True
    return trainer._loss_dev[:2]


def pick_main(trainer):
    """The double-Q coin flip, consuming the python RNG exactly like sqn_gru.py:207-216."""
    return 0 if random.uniform(0, 1) <= 0.5 else 1


def q_train_step(trainer, hp, s, a, r, s_next, true_len, true_next_len, is_end, main=None) -> torch.Tensor:
    """SQN_trainer / SMORL_trainer.train_step; returns device tensor [sup_loss, q_loss]."""
    B = int(s.shape[0])
    if main is None:
        main = pick_main(trainer)
    trainer.last_main = main + 1
    if trainer._shard is not None:
        return _sharded_step(trainer, hp, main, s, a, true_len, r, s_next, true_next_len, is_end)
    eng = trainer._ready(B)
    ds, dsn, da, dln, dnl, dr, de = trainer._stager.stage(s, a, true_len, r, s_next, true_next_len, is_end)
    batch = eng._batch(B, ds, da, dln, dr, dsn, dnl, de)
    eng.train_step_q(batch, hp, main, trainer._loss_dev)
    return trainer._loss_dev[:2]
