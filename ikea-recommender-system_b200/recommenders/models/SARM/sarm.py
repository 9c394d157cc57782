"""Drop-in for `recommenders/models/SARM/sarm.py` (reference :5-158): one net with five Q heads over the GRU state
(`MultiObjectiveQNetwork`), head 0 doubling as the supervised head -- executed natively on a B200
(`rec_train_step_sarm`).

Reference semantics kept on purpose ("quirky"): `gamma` is fixed to 0.99 in the trainer (:112); the
python RNG draw `random.randint(0, 4)` (:123) only selects tensors that never reach the loss, but it is consumed so
that a shared RNG stream stays aligned; the TD targets use max_a Q_i(s', a) of the SAME net, detached, with no is_end
masking (:133-135); loss = CE(Q_0(s, .), a) + mean_i mean_b (r + gamma max_a Q_i(s', a) - Q_i(s, a))^2 (:137-143)."""

import random

import torch

from ..._base import NativeSessionNet, NativeTrainerBase, make_hparams


class MultiObjectiveQNetwork(NativeSessionNet):
    def __init__(self, hidden_dim, item_num, state_size, action_dim, gru_layers, embedding_dim=50, random_embed_init=True,
                 train_pad_embed=True, use_packed_seq=False, padding_idx=None, name="QNetwork"):
        super().__init__()
        if not random_embed_init:
            raise NotImplementedError("TODO: Pretrained embeddings.")  # same as the reference
        self.random_embed_init = random_embed_init
        self.name = name
        self._build("sarm", hidden_dim, embedding_dim, item_num, state_size, action_dim, gru_layers, use_packed_seq,
                    train_pad_embed, padding_idx)


class SARM_trainer(NativeTrainerBase):
    def __init__(self, hidden_dim, embedding_dim, train_pad_embed, use_packed_seq, learning_rate, item_num, state_size,
                 action_dim, gru_layers, device, padding_idx=None, torch_rand_seed=118, python_rand_seed=999):
        self._seed(torch_rand_seed, python_rand_seed)
        self.network = MultiObjectiveQNetwork(hidden_dim=hidden_dim, item_num=item_num, state_size=state_size,
                                              action_dim=action_dim, gru_layers=gru_layers, embedding_dim=embedding_dim,
                                              train_pad_embed=train_pad_embed, use_packed_seq=use_packed_seq,
                                              padding_idx=padding_idx)
        self.gamma = 0.99
        self._setup([self.network], device, learning_rate)
        self.send_to_device()  # the reference moves the net in its constructor (:115)
        self.last_main_idx = None

    def _hp(self):
        return make_hparams(self.learning_rate, gamma=self.gamma, alpha=1.0, q_weights=(1.0, 0.0, 0.0))

    def train_step_async(self, s, a, r, s_next, true_len, true_next_len, is_end) -> torch.Tensor:
        """Device tensor [sup_loss, mean_i q_loss_i]."""
        self.last_main_idx = random.randint(0, 4)  # sarm.py:123 (the selected head's tensors never reach the loss)
        B = int(s.shape[0])
        eng = self._ready(B)
        ds, dsn, da, dln, dnl, dr, de = self._stager.stage(s, a, true_len, r, s_next, true_next_len, is_end)
        eng.train_step_sarm(eng._batch(B, ds, da, dln, dr, dsn, dnl, de), self._hp(), self._loss_dev)
        return self._loss_dev[:2]

    def train_step(self, s, a, r, s_next, true_len, true_next_len, is_end):
        out = self.train_step_async(s, a, r, s_next, true_len, true_next_len, is_end).tolist()
        return out[0], out[1]
