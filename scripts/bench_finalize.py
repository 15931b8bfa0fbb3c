#!/usr/bin/env python
"""bma_finalize_slots_kernel alone (the owner-side step of the S-sharded predictive) on one GPU: n_src source slots of
rows x (2C + 4) floats -> the finalised outputs of the owned rows.  CUDA events; a buffer larger than L2 per call."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortuna_b200 import ops  # noqa: E402

WANT = ("mean", "aleatoric_variance", "epistemic_variance", "entropy", "aleatoric_entropy", "epistemic_entropy", "log_prob",
        "mode", "confidence", "brier_terms")

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--C", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    for n_src in (2, 4, 8):
        rowf = ops.cls_slot_row_floats(a.C)
        bufs = [torch.rand((n_src, a.rows, rowf), device=dev) for _ in range(4)]  # rotate: no L2 reuse between calls
        t = torch.randint(0, a.C, (a.rows,), device=dev, dtype=torch.int32)
        out = None
        for i in range(5):
            out = ops.bma_finalize_cls_slots(bufs[i % 4], 64, a.C, WANT, out=out, targets=t)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for i in range(a.steps):
            ops.bma_finalize_cls_slots(bufs[i % 4], 64, a.C, WANT, out=out, targets=t)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / a.steps
        gb = (n_src * a.rows * rowf * 4 + 3 * a.rows * a.C * 4) / 1e9
        print(json.dumps({"n_src": n_src, "rows": a.rows, "C": a.C, "ms": ms, "GB": gb, "GBps": gb / ms * 1e3}))
