"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

from flux import specs, synthetic
from oracle import flux_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def golden(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def rel_l2(a, b):
    a = torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).detach().float().cpu()
    b = torch.as_tensor(np.asarray(b) if not isinstance(b, torch.Tensor) else b).detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def cosine(a, b):
    a = torch.as_tensor(a).detach().float().cpu().flatten()
    b = torch.as_tensor(b).detach().float().cpu().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-12)).item()


def small_configs():
    """The reduced configurations the golden fixtures were generated with (oracle/gen_golden.py)."""
    flow = json.loads(str(golden("flow_schnell.npz")["config"]))
    ae = json.loads(str(golden("ae_decode.npz")["config"]))
    te = golden("text_encoders.npz")
    return flow, ae, json.loads(str(te["t5_config"])), json.loads(str(te["clip_config"]))


def oracle_t5_config(t5c):
    return O.T5Config(**{k: v for k, v in t5c.items() if k in O.T5Config.__dataclass_fields__})


class FixedTokenizer:
    """Returns fixed token ids (the fixtures carry the ids; no tokenizer files exist offline)."""

    def __init__(self, ids):
        self.ids = torch.as_tensor(np.asarray(ids)).to(torch.int32)

    def encode(self, text, pad=True):
        return self.ids
