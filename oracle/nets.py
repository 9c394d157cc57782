"""Oracle networks: one spec-driven torch-CPU module for all four model families.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference has four near-identical nn.Modules; the oracle restates them as ONE
module, `SessionNet`, configured by a small spec.  Sub-module attribute names, the
construction order (=> torch RNG consumption order) and the init recipe follow the
reference so that `state_dict()` keys match and seeded init is bit-identical:

  GRU4Rec        recommenders/models/GRU4Rec/model.py:6-82      trunk "gru",        heads ["output"]
  BidirGRU4Rec   recommenders/models/BidirGRU4Rec/model.py:7-101 trunk "gru" (bidir), dropout, ["output"]
  SQN_Network    recommenders/models/SQN/sqn_gru.py:10-112       trunk "base_model", ["sup_head_output","q_head_output"]
  SMORL_GRU_Net  recommenders/models/SMORL/smorl_gru.py:14-139   trunk "base_model", ["sup_head_output","q_head_acc","q_head_div","q_head_nov"]
  Bidir-SQN      restatement for BASELINE cfg3: SQN heads on a bidirectional trunk
  MultiObjectiveQNetwork  recommenders/models/SARM/sarm.py:5-76  trunk "base_model", ModuleList q_heads[0..4]
"""

from __future__ import annotations

import torch
import torch.nn as nn
from torch.nn.utils.rnn import pack_padded_sequence


class SessionNet(nn.Module):
    """embedding -> (packed) GRU -> layer-0 final state(s) -> full-vocabulary linear heads."""

    def __init__(
        self,
        *,
        family: str,  # "gru4rec" | "bidir" | "sqn" | "smorl" | "bidir_sqn" | "sarm"
        hidden_dim: int,
        embedding_dim: int,
        item_num: int,
        state_size: int,
        action_dim: int,
        gru_layers: int = 1,
        dropout: float = 0.0,
        use_packed_seq: bool = True,
        train_pad_embed: bool = True,
        padding_idx=None,
    ):
        super().__init__()
        self.family = family
        self.hidden_dim = hidden_dim
        self.embedding_dim = embedding_dim
        self.item_num = int(item_num)
        self.state_size = state_size
        self.action_dim = action_dim
        self.gru_layers = gru_layers
        self.use_packed_seq = use_packed_seq
        self.bidirectional = family in ("bidir", "bidir_sqn")
        rl_family = family in ("sqn", "smorl", "bidir_sqn", "sarm")

        pad = self.item_num if padding_idx is None else padding_idx
        # SQN/SMORL force a trainable pad row under packing (sqn_gru.py:46-47, smorl_gru.py:51-52);
        # GRU4Rec/Bidir do not (GRU4Rec/model.py:35-47).
        if rl_family and use_packed_seq:
            train_pad_embed = True
        self.train_pad_embed = train_pad_embed

        # nn.Embedding default-inits with N(0,1) (consumes RNG) before the N(0, 0.01) re-init.
        self.embedding = nn.Embedding(self.item_num + 1, embedding_dim,
                                      padding_idx=None if train_pad_embed else pad)
        self.embedding.weight.data.normal_(mean=0, std=0.01)
        if not train_pad_embed:
            with torch.no_grad():
                self.embedding.weight[pad] = torch.zeros(embedding_dim)

        trunk = nn.GRU(input_size=embedding_dim, hidden_size=hidden_dim, num_layers=gru_layers,
                       bias=True, batch_first=True, bidirectional=self.bidirectional)
        self.trunk_name = "base_model" if rl_family else "gru"
        setattr(self, self.trunk_name, trunk)

        if family == "bidir":
            self.dropout = nn.Dropout(p=dropout)  # BidirGRU4Rec/model.py:60

        d = hidden_dim * (2 if self.bidirectional else 1)
        self.head_names = {
            "gru4rec": ["output"],
            "bidir": ["output"],
            "sqn": ["sup_head_output", "q_head_output"],
            "bidir_sqn": ["sup_head_output", "q_head_output"],
            "smorl": ["sup_head_output", "q_head_acc", "q_head_div", "q_head_nov"],
            "sarm": [],
        }[family]
        for name in self.head_names:
            setattr(self, name, nn.Linear(in_features=d, out_features=action_dim))
        if family == "sarm":  # sarm.py:58-60
            self.q_heads = nn.ModuleList([nn.Linear(in_features=d, out_features=action_dim) for _ in range(5)])

    # -- pieces ------------------------------------------------------------------------------
    def final_state(self, s, lengths):
        """[B, D] layer-0 final hidden state (concat fwd|bwd when bidirectional)."""
        seq = self.embedding(s)
        if self.use_packed_seq:
            seq = pack_padded_sequence(seq, lengths=lengths, batch_first=True, enforce_sorted=False)
        _, h = getattr(self, self.trunk_name)(seq)
        if self.bidirectional:
            return torch.cat([h[0, :, :], h[1, :, :]], dim=1)  # BidirGRU4Rec/model.py:90
        return h[0, :, :]  # layer 0 even when gru_layers > 1 (GRU4Rec/model.py:77, sqn_gru.py:104)

    def forward(self, s, lengths):
        h = self.final_state(s, lengths)
        if self.family == "sarm":
            return [head(h) for head in self.q_heads]  # sarm.py:74-75
        if self.family == "bidir":
            h = self.dropout(h)
        outs = [getattr(self, n)(h) for n in self.head_names]
        if len(outs) == 1:
            return outs[0]
        if len(outs) == 2:
            return outs[0], outs[1]
        return outs[0], torch.stack(outs[1:], dim=1)  # smorl_gru.py:137 -> [B, 3, V]


def _mk(family, **kw):
    return SessionNet(family=family, **kw)


def make_gru4rec(**kw):
    return _mk("gru4rec", **kw)


def make_bidir_gru4rec(**kw):
    return _mk("bidir", **kw)


def make_sqn(**kw):
    return _mk("sqn", **kw)


def make_smorl(**kw):
    return _mk("smorl", **kw)


def make_bidir_sqn(**kw):
    return _mk("bidir_sqn", **kw)


def make_sarm(**kw):
    return _mk("sarm", **kw)
