"""Synthetic parameter pytrees in Fortuna's layout for the tests and the benchmark."""
import numpy as np


def ragged_tree(seed=0, scale=0.01):
    """Odd sizes, 1-element leaves, unaligned halves: exercises every branch of the leaf walk."""
    r = np.random.default_rng(seed)
    f = lambda *s: (scale * r.standard_normal(s)).astype(np.float32)
    return {
        "model": {
            "params": {
                "dfe_subnet": {
                    "hidden1": {"bias": f(30), "kernel": f(2, 30)},
                    "hidden2": {"bias": f(7), "kernel": f(30, 7)},
                    "scalar": {"t": f(1)},
                    "odd": {"kernel": f(5, 5, 1, 6)},
                    "big": {"bias": f(3000), "kernel": f(64, 129)},
                    "wide": {"kernel": f(96, 128)},
                },
                "output_subnet": {"last": {"bias": f(2), "kernel": f(7, 2)}},
            }
        }
    }


def mlp_two_moons_tree(seed=0):
    """fortuna.model.MLP(2 -> 30 -> 30 -> 2): N = 1082 (BASELINE config 1)."""
    r = np.random.default_rng(seed)
    f = lambda *s: (0.1 * r.standard_normal(s)).astype(np.float32)
    return {
        "model": {
            "params": {
                "dfe_subnet": {
                    "hidden1": {"bias": f(30), "kernel": f(2, 30)},
                    "hidden2": {"bias": f(30), "kernel": f(30, 30)},
                },
                "output_subnet": {"last": {"bias": f(2), "kernel": f(30, 2)}},
            }
        }
    }


def lenet5_tree(seed=0):
    """fortuna.model.LeNet5 for MNIST: 156 + 2416 + 30840 + 10164 + 850 = 44426 params (config 2)."""
    r = np.random.default_rng(seed)
    f = lambda *s: (0.05 * r.standard_normal(s)).astype(np.float32)
    return {
        "model": {
            "params": {
                "dfe_subnet": {
                    "conv1": {"bias": f(6), "kernel": f(5, 5, 1, 6)},
                    "conv2": {"bias": f(16), "kernel": f(5, 5, 6, 16)},
                    "fc1": {"bias": f(120), "kernel": f(256, 120)},
                    "fc2": {"bias": f(84), "kernel": f(120, 84)},
                },
                "output_subnet": {"last": {"bias": f(10), "kernel": f(84, 10)}},
            }
        }
    }


from fortuna_b200.utils.tree_shapes import bert_base_shapes  # noqa: E402,F401  (the benchmark uses it too)
