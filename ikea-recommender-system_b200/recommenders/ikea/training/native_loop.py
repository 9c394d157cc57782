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


def eval_points(n_batches, eval_at):
    """Batch counts (1-based) after which the reference evaluates (trainSQN.py:160-161, :262)."""
    return sorted({int(n_batches * p) for p in eval_at})


def train_native(trainer, train_buffer, val_set, epochs, batch_size, val_batch_size, padding_pos, diversity_embedding,
                 unpopular_actions_set, eval_at=(0.5, 1.0), head_idx=0, topk_hr_ndcg=(5, 10, 20), topk_cov=(1, 5, 10),
                 topk_div=1, topk_nov=1, nov_rew_sig=1, input_tokenizer=None, output_tokenizer=None, generator=None,
                 out_dir=None, best_model_metric="val_hr", drop_last=False, log=None):
    """Returns the history: one dict per evaluation point with the reference's quantities
    (train_sup_loss, train_q_loss, train_hr/ndcg/reps, train_div_rew, train_nov_rew, train_cov and, per twin,
    val_loss/hr/ndcg/cov/r_div/r_nov/reps).  `train_buffer`: DeviceReplayBuffer on the trainer's device; `val_set`:
    DeviceEvaluationDataset on it (or any iterable of `(s, a, s_len)` batches)."""
    nets = _twins(trainer)
    m1 = nets[0]
    trainer.send_to_device()
    dev = m1._param_device()
    eng = trainer._ready(batch_size)
    bounds = train_buffer.batch_bounds(len(train_buffer), batch_size, drop_last)
    n_batches = len(bounds)
    points = set(eval_points(n_batches, eval_at))
    topk_hr_ndcg, topk_cov = list(topk_hr_ndcg), list(topk_cov)
    opts, kmax, keep = EP._opts(m1, dev, head_idx, topk_hr_ndcg, topk_div, topk_nov, topk_cov, nov_rew_sig, padding_pos,
                                diversity_embedding, unpopular_actions_set, input_tokenizer, output_tokenizer)
    unpop_np = keep[1].cpu().numpy()
    nk = len(topk_hr_ndcg)
    history, best = [], 0.0

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
            rec = dict(epoch=epoch, batch=n_batch + 1,
                       train_sup_loss=float(ls[0] / batch_counter), train_q_loss=float(ls[1] / batch_counter),
                       train_hr=rd["hits"][:nk] / n_samples, train_ndcg=rd["ndcg"][:nk] / n_samples,
                       train_reps=rd["reps"][:nk] / n_samples, train_div_rew=float(rd["div_sum"] / n_samples),
                       train_nov_rew=float(rd["nov_sum"] / n_samples),
                       train_cov=EP._coverage(rd["cov_bits"], topk_cov, unpop_np, m1.action_dim,
                                              len(unpopular_actions_set)))
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
                if out_dir is not None:
                    score = rec[f"{best_model_metric}{sfx}"]
                    score = float(np.asarray(score).reshape(-1)[-1])  # the largest k, like the reference's metric pick
                    if score > best:
                        best = score
                        os.makedirs(out_dir, exist_ok=True)
                        save_best_checkpoint(os.path.join(out_dir, "best_model.pt"), epoch, net, idx)
            trainer.set_train()
            history.append(rec)
            if log is not None:
                log(rec)
    return history
