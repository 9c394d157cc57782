"""Drop-in for `recommenders/models/BidirGRU4Rec/model.py` (reference :7-194): bidirectional GRU
trunk, concat(h_fwd, h_bwd) -> Dropout -> Linear(2H, V); supervised CE trainer."""

import torch

from .._native_models import host_path, supervised_train_step, supervised_train_step_host
from ..._base import NativeSessionNet, NativeTrainerBase


class BidirGRU4Rec(NativeSessionNet):
    def __init__(self, hidden_size, embedding_dim, item_num, state_size, action_dim, gru_layers=1, dropout=0,
                 use_packed_seq=True, train_pad_embed=True, padding_idx=None):
        super().__init__()
        self.layers = gru_layers
        self._build("bidir", hidden_size, embedding_dim, item_num, state_size, action_dim, gru_layers,
                    use_packed_seq, train_pad_embed, padding_idx, dropout=dropout)


class BidirGRU4Rec_trainer(NativeTrainerBase):
    def __init__(self, hidden_dim, embedding_dim, gru_layers, dropout, train_pad_embed, use_packed_seq,
                 learning_rate, item_num, state_size, action_dim, device, padding_idx=None, torch_rand_seed=118,
                 python_rand_seed=999):
        self._seed(torch_rand_seed, python_rand_seed)
        self._dropout_seed = int(torch_rand_seed)   # seed of the device-side dropout mask stream
        self.dropout_mask_override = None           # optional uint8 [B, 2H] keep mask for the next steps (tests)
        self.gru_model = BidirGRU4Rec(hidden_size=hidden_dim, embedding_dim=embedding_dim,
                                      train_pad_embed=train_pad_embed, use_packed_seq=use_packed_seq,
                                      item_num=item_num, state_size=state_size, action_dim=action_dim,
                                      gru_layers=gru_layers, dropout=dropout, padding_idx=padding_idx)
        self._setup([self.gru_model], device, learning_rate)

    def train_step(self, s, a, true_len):
        if host_path(self, s):
            return supervised_train_step_host(self, s, a, true_len)
        return supervised_train_step(self, s, a, true_len).item()

    def train_step_async(self, s, a, true_len) -> torch.Tensor:
        return supervised_train_step(self, s, a, true_len)
