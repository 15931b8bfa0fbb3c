"""Minimal array-backed loaders with the reference's constructor / conversion names
(fortuna/data/loader/array_loaders.py, fortuna/data/loader/base.py:31-533): enough of the loader protocol for
``ProbClassifier.train`` and ``predictive.*`` - iteration over (inputs, targets) / inputs / targets batches.
Host iteration is outside the hot path; batches are numpy arrays."""
from typing import Iterator, Optional, Tuple

import numpy as np


def _batches(n: int, batch_size: Optional[int], shuffle: bool, rng):
    idx = np.arange(n)
    if shuffle:
        rng.shuffle(idx)
    bs = n if not batch_size else int(batch_size)
    for i in range(0, n, bs):
        yield idx[i : i + bs]


class InputsLoader:
    def __init__(self, inputs: np.ndarray, batch_size: Optional[int] = None):
        self._inputs, self._batch_size = np.asarray(inputs), batch_size

    @classmethod
    def from_array_inputs(cls, inputs, batch_size: Optional[int] = None, shuffle: bool = False, prefetch: bool = False):
        return cls(inputs, batch_size)

    def to_array_inputs(self) -> np.ndarray:
        return self._inputs

    @property
    def size(self) -> int:
        return len(self._inputs)

    def __iter__(self) -> Iterator[np.ndarray]:
        for ix in _batches(len(self._inputs), self._batch_size, False, None):
            yield self._inputs[ix]


class TargetsLoader:
    def __init__(self, targets: np.ndarray, batch_size: Optional[int] = None):
        self._targets, self._batch_size = np.asarray(targets), batch_size

    @classmethod
    def from_array_targets(cls, targets, batch_size: Optional[int] = None, shuffle: bool = False, prefetch: bool = False):
        return cls(targets, batch_size)

    def to_array_targets(self) -> np.ndarray:
        return self._targets

    def __iter__(self) -> Iterator[np.ndarray]:
        for ix in _batches(len(self._targets), self._batch_size, False, None):
            yield self._targets[ix]


class DataLoader:
    def __init__(self, inputs, targets, batch_size: Optional[int] = None, shuffle: bool = False, seed: int = 0):
        self._inputs, self._targets = np.asarray(inputs), np.asarray(targets)
        if len(self._inputs) != len(self._targets):
            raise ValueError("inputs and targets must have the same number of data points")
        self._batch_size, self._shuffle = batch_size, shuffle
        self._rng = np.random.default_rng(seed)

    @classmethod
    def from_array_data(cls, data: Tuple, batch_size: Optional[int] = None, shuffle: bool = False, prefetch: bool = False):
        return cls(data[0], data[1], batch_size, shuffle)

    def to_inputs_loader(self) -> InputsLoader:
        return InputsLoader(self._inputs, self._batch_size)

    def to_targets_loader(self) -> TargetsLoader:
        return TargetsLoader(self._targets, self._batch_size)

    def to_array_inputs(self) -> np.ndarray:
        return self._inputs

    def to_array_targets(self) -> np.ndarray:
        return self._targets

    @property
    def size(self) -> int:
        return len(self._inputs)

    def __len__(self) -> int:
        bs = self._batch_size or len(self._inputs)
        return (len(self._inputs) + bs - 1) // bs

    def __iter__(self):
        for ix in _batches(len(self._inputs), self._batch_size, self._shuffle, self._rng):
            yield self._inputs[ix], self._targets[ix]
