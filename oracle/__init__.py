"""CPU oracle for the GRU4Rec / BidirGRU4Rec / SQN / SMORL hot path.

TEST INFRASTRUCTURE ONLY.  This package is a torch-CPU (fp32) restatement of the
reference's algorithm (adam-walsh-data/IKEA-Recommender-System, `recommenders/`).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it, and there only as the checker / the timed
CPU baseline -- never as the product path.  The product
(`ikea-recommender-system_b200/`) never imports anything from here and fails loudly
when its CUDA library is missing.

Parity pinning (see `oracle/validate_vs_reference.py`, run in the build container
where `/root/reference` exists): every net and trainer here is bit-identical to the
reference classes on shared seeds (same torch build), and every helper reproduces
the known answers of the reference's own unit tests (`test/test_tensor_operations.py`,
`test/test_evaluation.py:155-267`, `test/test_coverage.py`, `test/test_novelty.py`,
`test/test_repetions.py`, `test/test_diversity.py:5-19`).  The fixtures produced by
`oracle/make_golden.py` from the *real* reference are committed under `tests/golden/`.

Restated (not runnable in the reference at HEAD):
  * SMORL train step with the third (novelty) reward column restored,
  * SQN over a bidirectional GRU trunk (BASELINE cfg3),
  * tie-stable top-k (score desc, id asc) -- `torch.topk` order among ties is unspecified.
"""

from .nets import SessionNet, make_gru4rec, make_bidir_gru4rec, make_sqn, make_smorl, make_bidir_sqn, make_sarm  # noqa: F401
from .trainers import GRUTrainer, SQNTrainer, SMORLTrainer, SARMTrainer  # noqa: F401
from .evalproto import (  # noqa: F401
    stable_topk,
    evaluate,
    update_train_metrics,
    diversity_rewards,
    novelty_rewards,
    hits_and_ndcg,
    repetitions,
    coverage_update,
    coverage_result,
)
