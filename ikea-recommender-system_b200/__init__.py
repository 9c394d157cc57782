"""B200-native GRU4Rec / BidirGRU4Rec / SQN / SMORL train-step + full-catalog top-k evaluation.

Importing this package dlopens the in-tree CUDA library (`librecsys_b200.so`, sm_100a); a missing
library is an ImportError-time RuntimeError -- there is no eager/CPU fallback anywhere in the product.

Public surface (mirrors the reference `recommenders` package for the hot path):
    recommenders.models.GRU4Rec.model        GRU4Rec, GRU4Rec_trainer
    recommenders.models.BidirGRU4Rec.model   BidirGRU4Rec, BidirGRU4Rec_trainer
    recommenders.models.SQN.sqn_gru          SQN_Network, SQN_trainer
    recommenders.models.SMORL.smorl_gru      SMORL_GRU_Net, SMORL_trainer
    recommenders.models.SARM.sarm            MultiObjectiveQNetwork, SARM_trainer
    recommenders.evaluate.eval_protocol      evaluate, update_train_metrics, get_preds
    recommenders.data_utils.replay_buffer    DeviceReplayBuffer, DeviceEvaluationDataset (device-resident sampling)
    recommenders.ikea.training.native_loop   train_native (the train_SQN / train_SMORL loop without per-batch host trips)
"""

from . import _native

LIB = _native.load_library()  # raises if the CUDA library has not been built

from .engine import Engine, EvalAccumulators, NetTensors  # noqa: E402,F401
from .recommenders.models.GRU4Rec.model import GRU4Rec, GRU4Rec_trainer  # noqa: E402,F401
from .recommenders.models.BidirGRU4Rec.model import BidirGRU4Rec, BidirGRU4Rec_trainer  # noqa: E402,F401
from .recommenders.models.SQN.sqn_gru import SQN_Network, SQN_trainer  # noqa: E402,F401
from .recommenders.models.SMORL.smorl_gru import SMORL_GRU_Net, SMORL_trainer  # noqa: E402,F401
from .recommenders.models.SARM.sarm import MultiObjectiveQNetwork, SARM_trainer  # noqa: E402,F401
from .recommenders.evaluate.eval_protocol import evaluate, update_train_metrics, get_preds  # noqa: E402,F401
from .recommenders.data_utils.replay_buffer import DeviceReplayBuffer, DeviceEvaluationDataset  # noqa: E402,F401
from .recommenders.ikea.training.native_loop import train_native  # noqa: E402,F401

__all__ = ["Engine", "GRU4Rec", "GRU4Rec_trainer", "BidirGRU4Rec", "BidirGRU4Rec_trainer", "SQN_Network",
           "SQN_trainer", "SMORL_GRU_Net", "SMORL_trainer", "MultiObjectiveQNetwork", "SARM_trainer", "evaluate", "update_train_metrics", "get_preds",
           "DeviceReplayBuffer", "DeviceEvaluationDataset", "train_native"]
