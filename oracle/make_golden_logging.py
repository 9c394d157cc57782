"""Pins the logging-key restatement of the native loop (`native_loop.logging_dict_train`) against the UNMODIFIED
reference `get_logging_dict_train` (recommenders/utils/logging_SMORL.py:1-71): calls it in this container on fixed inputs
for both twins' prefixes and writes tests/golden/logging_train.json (inputs + the reference's dicts).

    python -m oracle.make_golden_logging
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def inputs():
    return dict(train_sup_loss=6.25, train_q_loss=0.125, val_loss=6.5, topk_hr_ndcg=[5, 10, 20],
                train_hr=[0.01, 0.02, 0.04], train_ndcg=[0.005, 0.0075, 0.0125], val_hr=[0.03, 0.05, 0.07],
                val_ndcg=[0.0125, 0.025, 0.03125], train_coverage_res={1: [0.1, 0.2], 5: [0.3, 0.4], 10: [0.5, 0.6]},
                val_coverage_res={1: [0.15, 0.25], 5: [0.35, 0.45], 10: [0.55, 0.65]}, topk_cov=[1, 5, 10],
                train_nov_rew=0.25, train_div_rew=0.75, val_nov_rew=0.375, val_div_rew=0.875,
                train_reps=[0.1, 0.2, 0.3], val_reps=[0.4, 0.5, 0.6])


def main():
    sys.path.insert(0, "/root/reference")
    from recommenders.utils.logging_SMORL import get_logging_dict_train
    kw = inputs()
    out = {"inputs": {k: ({str(a): b for a, b in v.items()} if isinstance(v, dict) else v) for k, v in kw.items()},
           "first": get_logging_dict_train(**kw, q_included=True, prefix=""),
           "second": get_logging_dict_train(**kw, q_included=True, prefix="Sec_")}
    with open(os.path.join(GOLD, "logging_train.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    msg = f"logging_train: reference get_logging_dict_train stored ({len(out['first'])} + {len(out['second'])} keys)"
    with open(os.path.join(GOLD, "VALIDATION.txt"), "a") as f:
        f.write(msg + "\n")
    print(msg)


if __name__ == "__main__":
    main()
