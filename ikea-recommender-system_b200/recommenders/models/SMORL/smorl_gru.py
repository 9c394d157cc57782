"""Drop-in for `recommenders/models/SMORL/smorl_gru.py` (reference :14-355): twin nets with a
supervised head and three Q heads (accuracy / diversity / novelty), scalarised double-Q with weights
`q_weights`, online diversity + novelty rewards, loss = sup + alpha * q.

The reference's train_step raises at HEAD (two reward columns for three heads, :311 vs :137); like the
oracle this class restores the novelty column (`topk_nov`, `nov_rew_sig`, commented out at :219-220)."""

import numpy as np
import torch

from .._native_models import host_path, q_train_step, q_train_step_host
from ..._base import NativeSessionNet, NativeTrainerBase, make_hparams
from ...evaluate.eval_protocol import _as_div_table, _unpopular_bitmap, _token_lut


class SMORL_GRU_Net(NativeSessionNet):
    def __init__(self, hidden_dim, embedding_dim, item_num, state_size, action_dim, q_weights, gamma, gru_layers=1,
                 random_embed_init=True, train_pad_embed=True, use_packed_seq=False, padding_idx=None,
                 name="SMORL_GRU_Network"):
        super().__init__()
        if not random_embed_init:
            raise NotImplementedError("TODO: Pretrained embedings.")
        self.random_embed_init = random_embed_init
        self.gamma = gamma
        self.q_weights = q_weights
        self.name = name
        self._build("smorl", hidden_dim, embedding_dim, item_num, state_size, action_dim, gru_layers,
                    use_packed_seq, train_pad_embed, padding_idx)


class SMORL_trainer(NativeTrainerBase):
    def __init__(self, hidden_dim, embedding_dim, padding_pos, train_pad_embed, use_packed_seq, learning_rate,
                 item_num, state_size, action_dim, gamma, gru_layers, q_weights, alpha, div_embedding,
                 unpopular_actions_set, topk_div, device, input_tokenizer=None, output_tokenizer=None,
                 padding_idx=None, torch_rand_seed=118, python_rand_seed=999, name_1="SMORL_1", name_2="SMORL_2",
                 topk_nov=1, nov_rew_sig=1.0):
        self._seed(torch_rand_seed, python_rand_seed)
        kw = dict(hidden_dim=hidden_dim, item_num=item_num, state_size=state_size, action_dim=action_dim,
                  gru_layers=gru_layers, gamma=gamma, q_weights=q_weights, embedding_dim=embedding_dim,
                  train_pad_embed=train_pad_embed, use_packed_seq=use_packed_seq, padding_idx=padding_idx)
        self.SMORL_1 = SMORL_GRU_Net(name=name_1, **kw)
        self.SMORL_2 = SMORL_GRU_Net(name=name_2, **kw)
        self.gamma = gamma
        self.alpha = alpha
        self.q_weights = q_weights
        self.padding_pos = padding_pos
        self.input_tokenizer = input_tokenizer
        self.output_tokenizer = output_tokenizer
        self.div_embedding = div_embedding
        self.topk_div = topk_div
        self.unpopular_actions_set = unpopular_actions_set
        self.topk_nov = topk_nov
        self.nov_rew_signal = nov_rew_sig
        self._setup([self.SMORL_1, self.SMORL_2], device, learning_rate)
        self.last_main = None
        self._hp_cache = None

    def _hp(self):
        eng_dev = self._nets[0]._param_device()
        if self._hp_cache is None or self._hp_cache[0] != eng_dev:
            w = [float(x) for x in torch.as_tensor(self.q_weights).reshape(-1).tolist()]
            hp = make_hparams(self.learning_rate, gamma=self.gamma, alpha=self.alpha, q_weights=w)
            div = _as_div_table(self.div_embedding, eng_dev)
            unpop = _unpopular_bitmap(self.unpopular_actions_set, self._nets[0].action_dim, eng_dev)
            lut = _token_lut(self.input_tokenizer, self.output_tokenizer, self._nets[0].action_dim, eng_dev)
            hp.div_emb, hp.div_dim = div.data_ptr(), int(div.shape[1])
            hp.topk_div, hp.topk_nov, hp.nov_reward = int(self.topk_div), int(self.topk_nov), float(self.nov_rew_signal)
            hp.unpopular = unpop.data_ptr()
            hp.out_to_in = None if lut is None else lut.data_ptr()
            hp.pad_pos_end = 1 if self.padding_pos == "end" else 0
            self._hp_cache = (eng_dev, hp, (div, unpop, lut))
        hp = self._hp_cache[1]
        hp.lr = self.learning_rate
        return hp

    def train_step(self, s, a, r_acc, s_next, true_len, true_next_len, is_end):
        """Scalarised double-Q step (reference :233-334, novelty column restored)."""
        self._ready(int(s.shape[0]))
        if host_path(self, s):
            return q_train_step_host(self, self._hp(), s, a, r_acc, s_next, true_len, true_next_len, is_end)
        out = q_train_step(self, self._hp(), s, a, r_acc, s_next, true_len, true_next_len, is_end).tolist()
        return out[0], out[1]

    def train_step_async(self, s, a, r_acc, s_next, true_len, true_next_len, is_end) -> torch.Tensor:
        self._ready(int(s.shape[0]))
        return q_train_step(self, self._hp(), s, a, r_acc, s_next, true_len, true_next_len, is_end)
