"""Native training-loop driver: the epoch / batch / evaluation structure of
`train_SQN` (ikea/training/trainSQN.py:168-428; `trainSMORL.py` has the same shape) on top of the device-resident
replay buffer and the fused kernels, without a host round trip per batch.

Reference loop, per batch: `train_step` (losses to the host) + `update_train_metrics` (second forward, five top-ks, a
B x V copy to the host, python sets for coverage).  Here, per batch: one `rec_gather_batch`, one graph replay of the
train step (losses accumulate in HBM) and one fused `rec_eval_batch` of `DQN_1` / `SMORL_1` on the same batch whose
hit / ndcg / repetition / reward sums and coverage bitmaps accumulate in device memory; the host reads them only at the
evaluation points `int(n_batches * p) for p in eval_at` (trainSQN.py:160-161), where both twins are evaluated on the
validation set and the best model is checkpointed in `SaveBestModel`'s format (utils/save_best_model.py:29-41).
Quirk q5 is kept: train metrics always score twin 1, after the optimizer step.
"""

import os

import numpy as np
import torch

from ...evaluate import eval_protocol as EP
from ....engine import EvalAccumulators


def _twins(trainer):
    nets = getattr(trainer, "_nets", None)
    if not nets or len(nets) != 2:
        raise TypeError("native_loop drives the twin-net trainers (SQN_trainer / SMORL_trainer)")
    return nets


def save_best_checkpoint(path, epoch, model, model_idx):
    """`SaveBestModel.__call__`'s payload (utils/save_best_model.py:29-41)."""
    torch.save({"epoch": epoch, "model_idx": model_idx, "hidden_dim": model.hidden_dim, "item_num": model.item_num,
                "action_dim": model.action_dim, "state_size": model.state_size, "embedding_dim": model.embedding_dim,
                "model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()}}, path)


# wandb / tensorboard key schema of one evaluation point (`get_logging_dict_train`, utils/logging_SMORL.py:1-71), as data:
# (key template, argument it reads, index into that argument).  "{p}" is the second twin's prefix, "{k}" a top-k value;
# a key without "Val" in it exists only in the first twin's (un-prefixed) record.
_LOG_SCALARS = [("Supervised Train Loss", "train_sup_loss"), ("Q-Modification-Signal", "train_q_loss"),
                ("{p} Supervised Val Loss", "val_loss")]
_LOG_PER_K = [("Train_HR@{k}", "train_hr"), ("Train_NDCG@{k}", "train_ndcg"), ("{p}Val_HR@{k}", "val_hr"),
              ("{p}Val_NDCG@{k}", "val_ndcg"), ("{p}Train_R@{k}", "train_reps"), ("{p}Val_R@{k}", "val_reps")]
_LOG_PER_COV_K = [("Train_NOV_CV@{k}", "train_coverage_res", 0), ("Train_DIV_CV@{k}", "train_coverage_res", 1),
                  ("{p}Val_NOV_CV@{k}", "val_coverage_res", 0), ("{p}Val_DIV_CV@{k}", "val_coverage_res", 1)]
_LOG_REWARDS = [("Train_Nov_Reward", "train_nov_rew"), ("Train_Div_Reward", "train_div_rew"),
                ("{p}Val_Nov_Reward", "val_nov_rew"), ("{p}Val_Div_Reward", "val_div_rew")]


def logging_dict_train(train_sup_loss, train_q_loss, val_loss, topk_hr_ndcg, train_hr, train_ndcg, val_hr, val_ndcg,
                       train_coverage_res, val_coverage_res, topk_cov, train_nov_rew, train_div_rew, val_nov_rew,
                       val_div_rew, train_reps, val_reps, q_included=True, prefix=""):
    """The record of one evaluation point under the reference's key names (schema tables above)."""
    src = locals()
    res = {}
    for tmpl, name in _LOG_SCALARS:
        if name != "train_q_loss" or q_included:
            res[tmpl.format(p=prefix)] = src[name]
    for i, k in enumerate(topk_hr_ndcg):
        for tmpl, name in _LOG_PER_K:
            res[tmpl.format(p=prefix, k=k)] = float(src[name][i])
    for k in topk_cov:
        for tmpl, name, j in _LOG_PER_COV_K:
            res[tmpl.format(p=prefix, k=k)] = float(src[name][k][j])
    for tmpl, name in _LOG_REWARDS:
        res[tmpl.format(p=prefix)] = float(src[name])
    return res if prefix == "" else {key: v for key, v in res.items() if "Val" in key}


def eval_points(n_batches, eval_at):
    """Batch counts (1-based) after which the reference evaluates (trainSQN.py:160-161, :262)."""
    return sorted({int(n_batches * p) for p in eval_at})


def train_native(trainer, train_buffer, val_set, epochs, batch_size, val_batch_size, padding_pos, diversity_embedding,
                 unpopular_actions_set, eval_at=(0.5, 1.0), head_idx=0, topk_hr_ndcg=(5, 10, 20), topk_cov=(1, 5, 10),
                 topk_div=1, topk_nov=1, nov_rew_sig=1, input_tokenizer=None, output_tokenizer=None, generator=None,
                 out_dir=None, best_model_metric=None, drop_last=False, log=None):
    """Returns the history: one dict per evaluation point with the reference's quantities -- raw (`train_sup_loss`,
    `train_q_loss`, `train_hr/ndcg/reps`, `train_div_rew`, `train_nov_rew`, `train_cov`, and per twin `val_loss/hr/ndcg/
    cov/r_div/r_nov/reps[_2]`) and as `logs`, the merged wandb records of both twins with the reference's key names
    (`{**epoch_log_res, **epoch_log_res_sec}`, trainSQN.py:337-402).  Like the reference, the train-side sums restart
    after every evaluation point (trainSQN.py:417-428), the better twin on `best_model_metric` (a logging key, default
    `Val_HR@<largest k>`) is offered to the best-model saver with `epoch = log_counter` (:382-400).
    `train_buffer`: DeviceReplayBuffer on the trainer's device; `val_set`: DeviceEvaluationDataset on it (or any
    re-iterable of `(s, a, s_len)` batches); `log`: optional callable receiving each `logs` dict (e.g. `wandb.log`)."""
    nets = _twins(trainer)
    m1 = nets[0]
    trainer.send_to_device()
    dev = m1._param_device()
    eng = trainer._ready(batch_size)
    bounds = train_buffer.batch_bounds(len(train_buffer), batch_size, drop_last)
    n_batches = len(bounds)
    points = set(eval_points(n_batches, eval_at))
    topk_hr_ndcg, topk_cov = list(topk_hr_ndcg), list(topk_cov)
    if best_model_metric is None:
        best_model_metric = f"Val_HR@{max(topk_hr_ndcg)}"
    opts, kmax, keep = EP._opts(m1, dev, head_idx, topk_hr_ndcg, topk_div, topk_nov, topk_cov, nov_rew_sig, padding_pos,
                                diversity_embedding, unpopular_actions_set, input_tokenizer, output_tokenizer)
    unpop_np = EP._unpopular_bitmap(unpopular_actions_set, m1.action_dim, dev, packed=True)  # packed bit words
    nk = len(topk_hr_ndcg)
    history, best, log_counter = [], 0.0, 0

    def val_batches():
        return val_set.batches(val_batch_size) if hasattr(val_set, "batches") else val_set

    for epoch in range(epochs):
        trainer.set_train()
        acc = EvalAccumulators(dev, m1.action_dim)
        loss_sum = torch.zeros(2, dtype=torch.float64, device=dev)
        n_samples = batch_counter = 0
        it = train_buffer.batches(eng, batch_size, shuffle=True, generator=generator, drop_last=drop_last)
        for n_batch, (s, a, r, s_next, s_len, s_next_len, is_end) in enumerate(it):
            B = int(s.shape[0])
            losses = trainer.train_step_async(s, a, r, s_next, s_len, s_next_len, is_end)
            loss_sum += losses[:2].to(torch.float64)
            # train metrics of twin 1 on the same (device-resident) batch, accumulated on the device (q5)
            e1 = m1._ready(B)
            e1.eval_batch(m1._net_id, e1._batch(B, s, a, s_len), opts, acc.struct)
            n_samples += B
            batch_counter += 1
            if n_batch + 1 not in points:
                continue
            rd = acc.read()
            ls = loss_sum.cpu().numpy()
            rec = dict(epoch=epoch, batch=n_batch + 1, log_counter=log_counter,
                       train_sup_loss=float(ls[0] / batch_counter), train_q_loss=float(ls[1] / batch_counter),
                       train_hr=rd["hits"][:nk] / n_samples, train_ndcg=rd["ndcg"][:nk] / n_samples,
                       train_reps=rd["reps"][:nk] / n_samples, train_div_rew=float(rd["div_sum"] / n_samples),
                       train_nov_rew=float(rd["nov_sum"] / n_samples),
                       train_cov=EP._coverage(rd["cov_bits"], topk_cov, unpop_np, m1.action_dim,
                                              len(unpopular_actions_set)))
            logs = {}
            for idx, net in enumerate(nets, start=1):
                v = EP.evaluate(val_batches(), net, dev, trainer.cross_entropy_loss, padding_pos, diversity_embedding,
                                unpopular_actions_set, head_idx=head_idx, topk_hr_ndcg=topk_hr_ndcg,
                                topk_to_consider_div=topk_div, topk_to_consider_nov=topk_nov,
                                topk_to_consider_cov=topk_cov, novelty_rew_signal=nov_rew_sig,
                                input_tokenizer=input_tokenizer, output_tokenizer=output_tokenizer)
                sfx = "" if idx == 1 else "_2"
                rec.update({f"val_loss{sfx}": float(v[0]), f"val_hr{sfx}": v[1], f"val_ndcg{sfx}": v[2],
                            f"val_cov{sfx}": v[3], f"val_r_div{sfx}": float(v[4]), f"val_r_nov{sfx}": float(v[5]),
                            f"val_reps{sfx}": v[6]})
                logs.update(logging_dict_train(rec["train_sup_loss"], rec["train_q_loss"], float(v[0]), topk_hr_ndcg,
                                               rec["train_hr"], rec["train_ndcg"], v[1], v[2], rec["train_cov"], v[3],
                                               topk_cov, rec["train_nov_rew"], rec["train_div_rew"], v[5], v[4],
                                               rec["train_reps"], v[6], q_included=True,
                                               prefix="" if idx == 1 else "Sec_"))
            if best_model_metric not in logs:
                raise KeyError(f"best_model_metric {best_model_metric!r} is not a logging key: {sorted(logs)}")
            first, second = logs[best_model_metric], logs["Sec_" + best_model_metric]
            best_idx = 2 if first < second else 1                      # trainSQN.py:382-392
            to_check = second if best_idx == 2 else first
            rec["best_model_idx"] = best_idx
            if out_dir is not None and to_check > best:                 # SaveBestModel.__call__ (maximisation)
                best = to_check
                os.makedirs(out_dir, exist_ok=True)
                save_best_checkpoint(os.path.join(out_dir, "best_model.pt"), log_counter, nets[best_idx - 1], best_idx)
            rec["logs"] = logs
            history.append(rec)
            if log is not None:
                log(logs)
            log_counter += 1
            # the reference restarts every train-side sum after an evaluation point (trainSQN.py:417-428)
            trainer.set_train()
            acc = EvalAccumulators(dev, m1.action_dim)
            loss_sum.zero_()
            n_samples = batch_counter = 0
    return history
