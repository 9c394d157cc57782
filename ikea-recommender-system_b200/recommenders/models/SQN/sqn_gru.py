"""Drop-in for `recommenders/models/SQN/sqn_gru.py` (reference :10-277): twin nets with a supervised
and a Q head, double-Q TD + cross-entropy, one Adam per net -- executed natively on a B200."""

import torch

from .._native_models import host_path, q_train_step, q_train_step_host
from ..._base import NativeSessionNet, NativeTrainerBase, make_hparams


class SQN_Network(NativeSessionNet):
    def __init__(self, hidden_dim, item_num, state_size, action_dim, gamma, gru_layers, embedding_dim=50,
                 random_embed_init=True, train_pad_embed=True, use_packed_seq=False, padding_idx=None,
                 name="DQNetwork", bidirectional=False):
        super().__init__()
        if not random_embed_init:
            raise NotImplementedError("TODO: Pretrained embedings.")  # same as the reference
        self.random_embed_init = random_embed_init
        self.gamma = gamma
        self.name = name
        self._build("bidir_sqn" if bidirectional else "sqn", hidden_dim, embedding_dim, item_num, state_size,
                    action_dim, gru_layers, use_packed_seq, train_pad_embed, padding_idx)


class SQN_trainer(NativeTrainerBase):
    def __init__(self, hidden_dim, embedding_dim, train_pad_embed, use_packed_seq, learning_rate, item_num,
                 state_size, action_dim, gamma, gru_layers, device, padding_idx=None, torch_rand_seed=118,
                 python_rand_seed=999, name_1="SQN_1", name_2="SQN_2", bidirectional=False):
        self._seed(torch_rand_seed, python_rand_seed)
        kw = dict(hidden_dim=hidden_dim, item_num=item_num, state_size=state_size, action_dim=action_dim,
                  gamma=gamma, gru_layers=gru_layers, embedding_dim=embedding_dim, train_pad_embed=train_pad_embed,
                  use_packed_seq=use_packed_seq, padding_idx=padding_idx, bidirectional=bidirectional)
        self.DQN_1 = SQN_Network(name=name_1, **kw)
        self.DQN_2 = SQN_Network(name=name_2, **kw)
        self.gamma = gamma
        self._setup([self.DQN_1, self.DQN_2], device, learning_rate)
        self.last_main = None

    def _hp(self):
        return make_hparams(self.learning_rate, gamma=self.gamma, alpha=1.0, q_weights=(1.0, 0.0, 0.0))

    def train_step(self, s, a, r, s_next, true_len, true_next_len, is_end):
        """Double-Q step (reference :183-254); returns (sup_loss, q_loss) as floats."""
        if host_path(self, s):
            return q_train_step_host(self, self._hp(), s, a, r, s_next, true_len, true_next_len, is_end)
        out = q_train_step(self, self._hp(), s, a, r, s_next, true_len, true_next_len, is_end).tolist()
        return out[0], out[1]

    def train_step_async(self, s, a, r, s_next, true_len, true_next_len, is_end) -> torch.Tensor:
        return q_train_step(self, self._hp(), s, a, r, s_next, true_len, true_next_len, is_end)
