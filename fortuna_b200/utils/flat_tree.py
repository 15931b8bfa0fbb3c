"""Parameter pytrees as ONE flat float32 CUDA vector plus a leaf table.

Fortuna's state pytree (nested FrozenDict ``{"model": {"params": {...}}}``) is kept
as nested dicts whose leaves are *views* into the flat vector, in jax tree-flatten order (dict keys
sorted at every level) - the order ``jax.random.split(key, num_leaves)`` hands out the per-leaf keys
(fortuna/utils/random.py:11-14).  The sampler kernel works on the flat vector and the leaf table.
"""

from typing import Any, Dict, List, Tuple

import numpy as np
import torch

from .. import ops

ALIGN = 4  # leaves start on 16-byte boundaries so the kernel can use 128-bit loads


def _flatten(tree, prefix=()) -> List[Tuple[Tuple[str, ...], Any]]:
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree.keys()):
            out.extend(_flatten(tree[k], prefix + (k,)))
        return out
    return [(prefix, tree)]


def _flatten_shapes(tree, prefix=()):
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree.keys()):
            out.extend(_flatten_shapes(tree[k], prefix + (k,)))
        return out
    return [(prefix, tuple(tree))]


def tree_paths(tree):
    return [p for p, _ in _flatten(tree)]


class FlatTree:
    """flat: float32 CUDA vector; leaf i occupies flat[offset_i : offset_i + len_i] with shape shapes[i]."""

    def __init__(self, flat: torch.Tensor, paths, shapes, offsets):
        self.flat = flat
        self.paths = list(paths)
        self.shapes = [tuple(s) for s in shapes]
        self.offsets = [int(o) for o in offsets]
        self.lens = [int(np.prod(s)) if len(s) else 1 for s in self.shapes]
        self._tables = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_tree(cls, tree: Dict, device="cuda") -> "FlatTree":
        """Copy a nested dict of arrays / tensors into a new flat vector (jax flatten order)."""
        items = _flatten(tree)
        if not items:
            raise ValueError("empty parameter tree")
        shapes, offsets, off = [], [], 0
        for _, leaf in items:
            shp = tuple(leaf.shape)
            n = int(np.prod(shp)) if len(shp) else 1
            if n == 0:
                raise ValueError("zero-size leaves are not supported")
            shapes.append(shp)
            offsets.append(off)
            off += (n + ALIGN - 1) // ALIGN * ALIGN
        flat = torch.zeros(off, dtype=torch.float32, device=device)
        ft = cls(flat, [p for p, _ in items], shapes, offsets)
        for (_, leaf), view in zip(items, ft.leaf_views()):
            src = leaf if isinstance(leaf, torch.Tensor) else torch.from_numpy(np.array(leaf))
            view.copy_(src.to(device=flat.device, dtype=torch.float32))
        return ft

    @classmethod
    def from_shapes(cls, shape_tree: Dict, device="cuda") -> "FlatTree":
        """Zero-initialised FlatTree from a nested dict whose leaves are shape tuples."""
        items = _flatten_shapes(shape_tree)
        shapes, offsets, off = [], [], 0
        for _, shp in items:
            n = int(np.prod(shp)) if len(shp) else 1
            shapes.append(tuple(shp))
            offsets.append(off)
            off += (n + ALIGN - 1) // ALIGN * ALIGN
        flat = torch.zeros(off, dtype=torch.float32, device=device)
        return cls(flat, [p for p, _ in items], shapes, offsets)

    def zeros_like(self) -> "FlatTree":
        ft = FlatTree(torch.zeros_like(self.flat), self.paths, self.shapes, self.offsets)
        ft._tables = self._tables
        return ft

    def empty_like(self) -> "FlatTree":
        ft = FlatTree(torch.empty_like(self.flat), self.paths, self.shapes, self.offsets)
        ft._tables = self._tables
        return ft

    def clone(self) -> "FlatTree":
        ft = FlatTree(self.flat.clone(), self.paths, self.shapes, self.offsets)
        ft._tables = self._tables
        return ft

    def same_layout(self, other: "FlatTree") -> bool:
        return self.paths == other.paths and self.shapes == other.shapes and self.offsets == other.offsets

    # ------------------------------------------------------------------ views
    def leaf_views(self) -> List[torch.Tensor]:
        return [self.flat[o : o + n].view(s) for o, n, s in zip(self.offsets, self.lens, self.shapes)]

    def tree(self) -> Dict:
        """Nested dict of views (Fortuna's params layout); writing through a view writes the flat vector."""
        root: Dict = {}
        for path, view in zip(self.paths, self.leaf_views()):
            d = root
            for k in path[:-1]:
                d = d.setdefault(k, {})
            if path:
                d[path[-1]] = view
            else:
                return view
        return root

    def to_numpy_tree(self) -> Dict:
        host = self.flat.detach().cpu().numpy()
        root: Dict = {}
        for path, o, n, s in zip(self.paths, self.offsets, self.lens, self.shapes):
            d = root
            for k in path[:-1]:
                d = d.setdefault(k, {})
            d[path[-1]] = host[o : o + n].reshape(s).copy()
        return root

    @property
    def n_params(self) -> int:
        return sum(self.lens)

    # ------------------------------------------------------------------ kernel tables
    def tables(self):
        """(leaves [L,2] int64, tiles [T,2] int32) device tables for fb200_sgmcmc_step_f32 (cached)."""
        if self._tables is None:
            self._tables = ops.build_leaf_tables(self.lens, self.offsets, self.flat.device)
        return self._tables


def as_flat_tree(x, like: "FlatTree" = None) -> "FlatTree":
    """Accept a FlatTree (zero-copy) or a nested dict (copied into `like`'s layout)."""
    if isinstance(x, FlatTree):
        if like is not None and not x.same_layout(like):
            raise ValueError("gradient and state trees have different structure")
        return x
    device = like.flat.device if like is not None else "cuda"
    ft = FlatTree.from_tree(x, device=device)
    if like is not None and not ft.same_layout(like):
        raise ValueError("gradient and state trees have different structure")
    return ft
