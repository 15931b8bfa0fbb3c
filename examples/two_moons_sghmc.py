#!/usr/bin/env python
"""BASELINE config 1: two-moons MLP ProbClassifier with an SGHMC posterior, predictive mean / entropy
(the reference's examples/two_moons_classification.pct.py with SGHMCPosteriorApproximator swapped in)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortuna_b200.data import DataLoader  # noqa: E402
from fortuna_b200.metric.classification import accuracy, expected_calibration_error  # noqa: E402
from fortuna_b200.prob_model import FitConfig, FitOptimizer, ProbClassifier, SGHMCPosteriorApproximator  # noqa: E402


class MLP(torch.nn.Module):
    """fortuna.model.MLP(output_dim, widths=(30, 30), activations=(tanh, tanh)): dfe_subnet + output_subnet."""

    def __init__(self, in_dim=2, widths=(30, 30), output_dim=2):
        super().__init__()
        layers, d = [], in_dim
        for w in widths:
            layers += [torch.nn.Linear(d, w), torch.nn.Tanh()]
            d = w
        self.dfe_subnet = torch.nn.Sequential(*layers)
        self.output_subnet = torch.nn.Linear(d, output_dim)

    def forward(self, x):
        return self.output_subnet(self.dfe_subnet(x))


def make_moons(n, noise, seed):
    r = np.random.default_rng(seed)
    n0 = n // 2
    t0, t1 = r.uniform(0, np.pi, n0), r.uniform(0, np.pi, n - n0)
    x = np.concatenate([np.stack([np.cos(t0), np.sin(t0)], 1), np.stack([1 - np.cos(t1), 0.5 - np.sin(t1)], 1)])
    y = np.concatenate([np.zeros(n0, np.int64), np.ones(n - n0, np.int64)])
    x = x + noise * r.standard_normal(x.shape)
    p = r.permutation(n)
    return x[p].astype(np.float32), y[p]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--step", type=float, default=1e-4)
    ap.add_argument("--epochs", type=int, default=300)
    ap.add_argument("--burnin", type=int, default=800)
    ap.add_argument("--n-samples", type=int, default=30)
    ap.add_argument("--thin", type=int, default=10)
    a = ap.parse_args()
    torch.manual_seed(0)
    train, test = make_moons(500, 0.07, 0), make_moons(500, 0.07, 2)
    train_loader = DataLoader.from_array_data(train, batch_size=128, shuffle=True)
    test_loader = DataLoader.from_array_data(test, batch_size=128)
    prob_model = ProbClassifier(
        model=MLP(), posterior_approximator=SGHMCPosteriorApproximator(
            n_samples=a.n_samples, n_thinning=a.thin, burnin_length=a.burnin, step_schedule=a.step))
    t0 = time.time()
    status = prob_model.train(train_loader, fit_config=FitConfig(optimizer=FitOptimizer(n_epochs=a.epochs)))
    torch.cuda.synchronize()
    dt = time.time() - t0
    steps = a.epochs * len(train_loader)
    inputs = test_loader.to_inputs_loader()
    stats = prob_model.predictive.statistics(test_loader, ("mean", "mode", "entropy", "epistemic_entropy", "log_prob"), 30)
    acc = float(accuracy(stats["mode"], test[1]))
    ece = float(expected_calibration_error(stats["mode"], stats["mean"], test[1]))
    # uncertainty far from the data should exceed uncertainty on it
    grid = np.stack(np.meshgrid(np.linspace(-4, 5, 30), np.linspace(-4, 4, 30)), -1).reshape(-1, 2).astype(np.float32)
    far = grid[(np.abs(grid[:, 0] - 0.5) > 3) | (np.abs(grid[:, 1]) > 3)]
    from fortuna_b200.data import InputsLoader

    ent_far = float(prob_model.predictive.epistemic_entropy(InputsLoader.from_array_inputs(far), 30).mean())
    ent_on = float(stats["epistemic_entropy"].mean())
    print(f"steps {steps} in {dt:.1f}s ({1e3 * dt / steps:.2f} ms/step) loss {status['fit_status']['loss'][-1]:.1f} "
          f"acc {acc:.3f} ece {ece:.3f} log_prob {float(stats['log_prob'].mean()):.3f} "
          f"epistemic entropy on-data {ent_on:.4f} far {ent_far:.4f}")


if __name__ == "__main__":
    main()
