"""Drop-in for `recommenders/models/GRU4Rec/model.py` (reference :6-173), executed natively on a B200.

Same class names, constructor keywords, attributes (`hidden_dim, item_num, action_dim, state_size,
embedding_dim`), `state_dict()` keys (`embedding.weight`, `gru.*_l0`, `output.*`) and
`train_step(s, a, true_len) -> float`.
"""

import torch

from .._native_models import host_path, supervised_train_step, supervised_train_step_host
from ..._base import NativeSessionNet, NativeTrainerBase


class GRU4Rec(NativeSessionNet):
    def __init__(self, hidden_size, embedding_dim, item_num, state_size, action_dim, gru_layers=1,
                 use_packed_seq=True, train_pad_embed=True, padding_idx=None):
        super().__init__()
        self.layers = gru_layers
        self._build("gru4rec", hidden_size, embedding_dim, item_num, state_size, action_dim, gru_layers,
                    use_packed_seq, train_pad_embed, padding_idx)


class GRU4Rec_trainer(NativeTrainerBase):
    def __init__(self, hidden_dim, embedding_dim, gru_layers, train_pad_embed, use_packed_seq, learning_rate,
                 item_num, state_size, action_dim, device, padding_idx=None, torch_rand_seed=118,
                 python_rand_seed=999):
        self._seed(torch_rand_seed, python_rand_seed)
        self.gru_model = GRU4Rec(hidden_size=hidden_dim, embedding_dim=embedding_dim, train_pad_embed=train_pad_embed,
                                 use_packed_seq=use_packed_seq, item_num=item_num, state_size=state_size,
                                 action_dim=action_dim, gru_layers=gru_layers, padding_idx=padding_idx)
        self._setup([self.gru_model], device, learning_rate)

    def train_step(self, s, a, true_len):
        """One supervised step (reference :129-155); returns the batch-mean CE loss as a float."""
        if host_path(self, s):
            return supervised_train_step_host(self, s, a, true_len)
        return supervised_train_step(self, s, a, true_len).item()

    def train_step_async(self, s, a, true_len) -> torch.Tensor:
        """Same step without the host sync: returns a 0-d device tensor."""
        return supervised_train_step(self, s, a, true_len)
