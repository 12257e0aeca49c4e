"""Shared helpers of the test-suite: golden fixtures, the oracle model, the product model built from the same
tensors.  (The oracle is imported here and in tests only - never by the product package.)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_cache = {}


def golden_npz(name):
    if name not in _cache:
        _cache[name] = dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))
    return _cache[name]


def golden_json(name):
    with open(os.path.join(GOLDEN, name), encoding="utf-8") as f:
        return json.load(f)


def oracle_model(name="ViT-B/16", seed=0):
    """fp32 CPU oracle CLIP with the seeded weights the goldens were generated with."""
    from oracle import clip_ref
    key = ("oracle", name, seed)
    if key not in _cache:
        _cache[key] = clip_ref.build_model(name, seed=seed)
    return _cache[key]


def oracle_state_dict(name="ViT-B/16", seed=0):
    return {k: v.clone() for k, v in oracle_model(name, seed).state_dict().items()}


def label_layout():
    lab = golden_json("labels.json")
    sizes = [len(lab["detector"])] + [len(lab["groups"][g]) for g in lab["group_order"]]
    split = [lab["n_interior"]] + [0] * len(lab["group_order"])
    return lab, sizes, split


def top5_sets(topk_idx_row, lab, g0=1):
    """engine top-k indices [G, k] -> {group: [labels]}"""
    out = {}
    for gi, g in enumerate(lab["group_order"]):
        names = lab["groups"][g]
        out[g] = [names[int(i)] for i in topk_idx_row[g0 + gi] if int(i) >= 0]
    return out
