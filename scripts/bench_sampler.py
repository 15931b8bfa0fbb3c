#!/usr/bin/env python
"""K1 measurement alone: SGHMC step on the BERT-base-shaped flat tree (c5), identity and RMSprop."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fortuna_b200.bench_extras import sampler_extras  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    print(json.dumps(sampler_extras(torch.device("cuda:0"), float(peaks["hbm_gbs"]), steps=a.steps, joint=True)))
