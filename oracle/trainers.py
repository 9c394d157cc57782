"""Oracle trainers: torch-CPU restatement of the reference `*_trainer.train_step`s.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  GRUTrainer    GRU4Rec_trainer       recommenders/models/GRU4Rec/model.py:85-155
                BidirGRU4Rec_trainer  recommenders/models/BidirGRU4Rec/model.py:104-176
  SQNTrainer    SQN_trainer           recommenders/models/SQN/sqn_gru.py:115-254
  SMORLTrainer  SMORL_trainer         recommenders/models/SMORL/smorl_gru.py:142-334  (restated: the
                reference stacks 2 reward columns for 3 Q-heads and raises at HEAD; here the novelty
                column is restored from novelty.py:12-47 -- SURVEY.md section 8c)

Quirks kept on purpose: (q1) the bootstrap net sees `s_next` with
`true_len`; (q2) SMORL's diversity reward indexes `s` with `true_next_len`; layer-0 state
only; the python `random` stream picks the main net.
"""

from __future__ import annotations

import random

import torch
import torch.nn as nn

from .nets import SessionNet
from .evalproto import diversity_rewards, novelty_rewards


def _seed(torch_rand_seed, python_rand_seed):
    torch.manual_seed(torch_rand_seed)
    random.seed(python_rand_seed)


class GRUTrainer:
    """Supervised next-item trainer (GRU4Rec / BidirGRU4Rec): forward -> CE -> Adam."""

    def __init__(self, *, family="gru4rec", hidden_dim, embedding_dim, gru_layers, train_pad_embed,
                 use_packed_seq, learning_rate, item_num, state_size, action_dim, device="cpu",
                 dropout=0.0, padding_idx=None, torch_rand_seed=118, python_rand_seed=999):
        _seed(torch_rand_seed, python_rand_seed)
        self.gru_model = SessionNet(family=family, hidden_dim=hidden_dim, embedding_dim=embedding_dim,
                                    item_num=item_num, state_size=state_size, action_dim=action_dim,
                                    gru_layers=gru_layers, dropout=dropout, use_packed_seq=use_packed_seq,
                                    train_pad_embed=train_pad_embed, padding_idx=padding_idx)
        self.device = device
        self.cross_entropy_loss = nn.CrossEntropyLoss(reduction="mean")
        self.optimizer = torch.optim.Adam(self.gru_model.parameters(), lr=learning_rate)

    def train_step(self, s, a, true_len):
        logits = self.gru_model(s, true_len)
        loss = self.cross_entropy_loss(logits, a)
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss.item()


class _TwinTrainer:
    """Two identical nets + two Adam optimizers; python RNG picks (main, boot) per step."""

    def _build(self, family, net_kw, learning_rate, torch_rand_seed, python_rand_seed):
        _seed(torch_rand_seed, python_rand_seed)
        self.net_1 = SessionNet(family=family, **net_kw)
        self.net_2 = SessionNet(family=family, **net_kw)
        self.cross_entropy_loss = nn.CrossEntropyLoss(reduction="mean")
        self.optimizer_1 = torch.optim.Adam(self.net_1.parameters(), lr=learning_rate)
        self.optimizer_2 = torch.optim.Adam(self.net_2.parameters(), lr=learning_rate)
        self.last_main = None

    def _pick(self):
        # sqn_gru.py:207-216 / smorl_gru.py:258-266
        if random.uniform(0, 1) <= 0.5:
            self.last_main = 1
            return self.net_1, self.net_2, self.optimizer_1
        self.last_main = 2
        return self.net_2, self.net_1, self.optimizer_2


class SQNTrainer(_TwinTrainer):
    def __init__(self, *, family="sqn", hidden_dim, embedding_dim, train_pad_embed, use_packed_seq,
                 learning_rate, item_num, state_size, action_dim, gamma, gru_layers, device="cpu",
                 padding_idx=None, torch_rand_seed=118, python_rand_seed=999):
        self._build(family, dict(hidden_dim=hidden_dim, embedding_dim=embedding_dim, item_num=item_num,
                                 state_size=state_size, action_dim=action_dim, gru_layers=gru_layers,
                                 use_packed_seq=use_packed_seq, train_pad_embed=train_pad_embed,
                                 padding_idx=padding_idx),
                    learning_rate, torch_rand_seed, python_rand_seed)
        self.gamma = gamma
        self.device = device

    @property
    def DQN_1(self):
        return self.net_1

    @property
    def DQN_2(self):
        return self.net_2

    def train_step(self, s, a, r, s_next, true_len, true_next_len, is_end):
        r = r.unsqueeze(1)
        main, boot, opt = self._pick()
        sup, q_all = main(s, true_len)
        q_sa = q_all.gather(1, a.unsqueeze(1))
        with torch.no_grad():
            _, q_next_main = main(s_next, true_next_len)
            a_star = torch.argmax(q_next_main, dim=1, keepdim=True)
            _, q_next_boot = boot(s_next, true_len)  # (q1) true_len, sqn_gru.py:231
            boot_val = q_next_boot.gather(1, a_star)
            boot_val[is_end] = 0.0
        q_loss = torch.mean((r + self.gamma * boot_val - q_sa) ** 2)
        sup_loss = self.cross_entropy_loss(sup, a)
        total = q_loss + sup_loss  # sqn_gru.py:245
        opt.zero_grad()
        total.backward()
        opt.step()
        return sup_loss.item(), q_loss.item()


class SMORLTrainer(_TwinTrainer):
    def __init__(self, *, hidden_dim, embedding_dim, padding_pos, train_pad_embed, use_packed_seq,
                 learning_rate, item_num, state_size, action_dim, gamma, gru_layers, q_weights, alpha,
                 div_embedding, unpopular_actions_set, topk_div, topk_nov=1, nov_rew_sig=1.0,
                 device="cpu", out_to_in=None, padding_idx=None, torch_rand_seed=118,
                 python_rand_seed=999):
        self._build("smorl", dict(hidden_dim=hidden_dim, embedding_dim=embedding_dim, item_num=item_num,
                                  state_size=state_size, action_dim=action_dim, gru_layers=gru_layers,
                                  use_packed_seq=use_packed_seq, train_pad_embed=train_pad_embed,
                                  padding_idx=padding_idx),
                    learning_rate, torch_rand_seed, python_rand_seed)
        self.gamma = gamma
        self.alpha = alpha
        self.q_weights = torch.as_tensor(q_weights, dtype=torch.float32)
        self.padding_pos = padding_pos
        self.div_embedding = div_embedding  # frozen nn.Embedding or [N+1, Ediv] tensor
        self.unpopular_actions_set = unpopular_actions_set
        self.topk_div = topk_div
        self.topk_nov = topk_nov
        self.nov_rew_sig = nov_rew_sig
        self.out_to_in = out_to_in
        self.device = device

    @property
    def SMORL_1(self):
        return self.net_1

    @property
    def SMORL_2(self):
        return self.net_2

    def train_step(self, s, a, r_acc, s_next, true_len, true_next_len, is_end):
        main, boot, opt = self._pick()
        sup, q_all = main(s, true_len)  # q_all [B, 3, V]
        sup_loss = self.cross_entropy_loss(sup, a)
        heads = q_all.size(1)
        pick = a.view(-1, 1, 1).expand(-1, heads, 1)
        q_sa = torch.gather(q_all, 2, pick).squeeze(2)  # tensor_operations.py:4-33
        w = self.q_weights
        with torch.no_grad():
            _, q_next_main = main(s_next, true_next_len)
            scalarised = torch.sum(q_next_main * w.view(1, -1, 1), dim=1)  # tensor_operations.py:50-70
            a_star = torch.argmax(scalarised, dim=1)  # tensor_operations.py:73-84
            _, q_next_boot = boot(s_next, true_len)  # (q1) smorl_gru.py:291
            boot_val = torch.gather(q_next_boot, 2, a_star.view(-1, 1, 1).expand(-1, heads, 1)).squeeze(2)
            boot_val[is_end, :] = 0.0
            r_div = diversity_rewards(s, sup, true_next_len, self.padding_pos, self.topk_div,
                                      self.div_embedding, out_to_in=self.out_to_in)  # (q2) smorl_gru.py:298-308
            r_nov = torch.as_tensor(
                novelty_rewards(sup, self.unpopular_actions_set, self.nov_rew_sig, self.topk_nov))
            r = torch.stack([r_acc.reshape(-1).to(torch.float32), r_div.to(torch.float32),
                             r_nov.to(torch.float32)], dim=1)
        per_head = (r + self.gamma * boot_val - q_sa) ** 2
        q_loss = torch.mean(torch.matmul(per_head, w))  # tensor_operations.py:36-47
        total = sup_loss + self.alpha * q_loss  # smorl_gru.py:325
        opt.zero_grad()
        total.backward()
        opt.step()
        return sup_loss.item(), q_loss.item()


class SARMTrainer:
    """SARM_trainer (recommenders/models/SARM/sarm.py:78-158): one net with five Q heads, head 0 doubles as the
    supervised head; gamma is fixed to 0.99 (:112); the `random.randint` draw (:123) only selects tensors that never
    reach the loss, but it is consumed like in the reference."""

    def __init__(self, *, hidden_dim, embedding_dim, train_pad_embed, use_packed_seq, learning_rate, item_num,
                 state_size, action_dim, gru_layers, device="cpu", padding_idx=None, torch_rand_seed=118,
                 python_rand_seed=999):
        _seed(torch_rand_seed, python_rand_seed)
        self.network = SessionNet(family="sarm", hidden_dim=hidden_dim, embedding_dim=embedding_dim, item_num=item_num,
                                  state_size=state_size, action_dim=action_dim, gru_layers=gru_layers,
                                  use_packed_seq=use_packed_seq, train_pad_embed=train_pad_embed, padding_idx=padding_idx)
        self.device = device
        self.gamma = 0.99
        self.cross_entropy_loss = nn.CrossEntropyLoss(reduction="mean")
        self.optimizer = torch.optim.Adam(self.network.parameters(), lr=learning_rate)
        self.last_main_idx = None

    def train_step(self, s, a, r, s_next, true_len, true_next_len, is_end):
        r = r.unsqueeze(1)
        outputs = self.network(s, true_len)
        self.last_main_idx = random.randint(0, 4)  # :123
        a_idx = a.unsqueeze(1)
        with torch.no_grad():
            outputs_next = self.network(s_next, true_next_len)
        q_losses = []
        for i in range(5):  # :133-135 (no is_end masking: the masked tensor of :131 is never used)
            nxt = outputs_next[i].gather(1, torch.argmax(outputs_next[i], dim=1).unsqueeze(1))
            q_losses.append(torch.mean((r + self.gamma * nxt - outputs[i].gather(1, a_idx)) ** 2))
        sup_loss = self.cross_entropy_loss(outputs[0], a)
        total = sup_loss + sum(q_losses) * (1.0 / len(q_losses))
        self.optimizer.zero_grad()
        total.backward()
        self.optimizer.step()
        return sup_loss.item(), sum(q.item() for q in q_losses) / len(q_losses)
