"""Minimal stand-in for the reference's yacs config (config/defaults.py:8-73): same attribute paths and
defaults, yaml merge, no yacs dependency."""
import copy

import yaml


class Node(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def _node(d):
    return Node({k: _node(v) if isinstance(v, dict) else copy.deepcopy(v) for k, v in d.items()})


DEFAULTS = {
    "OUTPUT_DIR": "output", "LOG_DIR": "log", "PRINT_FREQ": 100, "SEED": 0,
    "MODEL": {"NAME": "deeplabv3+", "BACKBONE": "mobilenet_v2", "PRETRAINED_WEIGHTS": ""},
    "CONTROLLER": {"NAME": "controller", "LOSS": "ppo", "PENALTY": 0.00001, "L": 2, "M": 6, "T": 2, "C": 2.5,
                   "NUM_MAGS": 10, "EXCLUDE_OPS_NUM": 0, "EXCLUDE_OPS": []},
    "DISCRIMINATOR": {"NAME": "momentum_feature"},
    "DATASET": {"ROOT": "./dataset", "NAME": "cifar10", "TRAINSET": "", "TESTSET": "",
                "DG": {"TRAIN": [1, 2, 3], "TEST": [4]}},
    "TRAIN": {"LR": 0.1, "WD": 0.0004, "BEGIN_EPOCH": 0, "WARMUP_EPOCH": 0, "END_EPOCH": 200, "BATCH_SIZE": 8,
              "SHUFFLE": True},
    "TEST": {"BATCH_SIZE": 8, "MODEL_DIR": ""},
}


def get_config(yaml_path=None, **overrides):
    cfg = _node(DEFAULTS)
    if yaml_path:
        with open(yaml_path) as f:
            _merge(cfg, yaml.safe_load(f) or {})
    _merge(cfg, overrides)
    return cfg


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = _node(v) if isinstance(v, dict) else v


def optic_search_config(backbone="resnet50"):
    """experiments/optic_sinkhorn/diversity.yaml with the north-star backbone."""
    return get_config(PRINT_FREQ=10, DATASET={"NAME": "optic", "DG": {"TRAIN": [1, 2, 3], "TEST": [4]}},
                      MODEL={"NAME": "deeplabv3+", "BACKBONE": backbone},
                      CONTROLLER={"LOSS": "ppo", "M": 6, "T": 2, "C": 2.5},
                      TRAIN={"LR": 0.001, "WD": 0.0, "BATCH_SIZE": 8, "WARMUP_EPOCH": 30, "END_EPOCH": 150})
