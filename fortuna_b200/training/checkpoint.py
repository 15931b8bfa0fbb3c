"""Checkpoints in the reference's on-disk format: the wire format and the
state-dict shape are Fortuna's, so its sample files load into the device sample bank (and the last-layer predictive) as
they are.  What a file's ``params`` tree looks like follows the MODEL definition, though: Fortuna's flax modules store
``.../kernel`` [D, C], a torch ``nn.Linear`` fitted here stores ``.../weight`` [C, D] - the bytes round-trip either way
(``SGMCMCPosterior.last_layer`` accepts both), but loading one framework's files into the other's backbone needs a
name / transpose map that is not provided.

The reference saves every state through ``flax.training.checkpoints.save_checkpoint`` / ``restore_checkpoint``
(fortuna/training/mixin.py:32-85): one file ``<dir>/<prefix><step>`` holding
``flax.serialization.msgpack_serialize(flax.serialization.to_state_dict(state))``.  flax 0.8.5 is not installable
here, so its published wire format is restated (flax/serialization.py at that version):

  * the state dict is a nested ``dict`` with ``str`` keys; tuples and lists become ``{"0": .., "1": ..}``; NamedTuples and
    struct dataclasses become ``{field: ..}``; ``None`` stays ``None``;
  * every array leaf is a msgpack ``ExtType(1, packb((shape, dtype.name, raw C-order bytes), use_bin_type=True))``;
    numpy scalars are ``ExtType(3, <same payload for a 0-d array>)``; python complex is ``ExtType(2, packb((re, im)))``;
  * arrays larger than 2**30 bytes are split: ``{"__msgpack_chunked_array__": True, "shape": {"0": ..}, "chunks":
    {"0": <flat chunk>, ..}}`` with at most ``2**30 // itemsize`` elements per chunk;
  * ``packb(..., default=ext_pack, strict_types=True)`` / ``unpackb(..., ext_hook=ext_unpack, raw=False)``.

Checkpoint files follow flax's naming: ``latest_checkpoint`` picks the file with the largest step under a natural
(numeric-aware) sort, ``save_checkpoint`` writes to ``<prefix>tmp`` and renames, then keeps the newest ``keep`` files.

Parity unpinned: no flax here and no checkpoint fixture in the reference's tests; the round trip, the layout of the
bytes (a hand-assembled msgpack document in tests/test_checkpoint.py) and the state-dict shape Fortuna expects
(``d["encoded_name"].values()`` -> state class name, fortuna/training/mixin.py:83) are what the tests hold.
"""

import os
import re
from typing import Any, Dict, Optional

import msgpack
import numpy as np

EXT_NDARRAY = 1
EXT_NATIVE_COMPLEX = 2
EXT_NPSCALAR = 3
MAX_CHUNK_SIZE = 2**30
_CHUNK_FLAG = "__msgpack_chunked_array__"


class InvalidCheckpointError(Exception):
    """A checkpoint at the same or a later step exists and ``overwrite`` is False (flax.errors.InvalidCheckpointError)."""


# ------------------------------------------------------------------------------------------------ state dicts
def to_state_dict(x: Any) -> Any:
    """flax.serialization.to_state_dict for the containers this package uses."""
    try:
        import torch

        if isinstance(x, torch.Tensor):
            t = x.detach()
            if t.dtype == torch.bfloat16:
                raise TypeError("bfloat16 leaves are stored as float32 in SG-MCMC states (SURVEY appendix D)")
            return t.cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    from ..utils.flat_tree import FlatTree

    if isinstance(x, FlatTree):
        return to_state_dict(x.to_numpy_tree())
    if x is None or isinstance(x, (bool, int, float, str, bytes, np.ndarray, np.generic)):
        return x
    if isinstance(x, dict):
        return {str(k): to_state_dict(v) for k, v in x.items()}
    if hasattr(x, "_fields") and isinstance(x, tuple):  # NamedTuple -> field names
        return {f: to_state_dict(getattr(x, f)) for f in x._fields}
    if isinstance(x, (list, tuple)):
        return {str(i): to_state_dict(v) for i, v in enumerate(x)}
    if hasattr(x, "__dataclass_fields__"):
        return {f: to_state_dict(getattr(x, f)) for f in x.__dataclass_fields__}
    raise TypeError(f"cannot serialise {type(x)!r}")


# ------------------------------------------------------------------------------------------------ msgpack layer
def _ndarray_to_bytes(arr: np.ndarray) -> bytes:
    if arr.dtype.hasobject or arr.dtype.isalignedstruct:
        raise ValueError("Object and structured dtypes not supported for serialization of ndarrays.")
    return msgpack.packb((tuple(int(s) for s in arr.shape), arr.dtype.name, arr.tobytes("C")), use_bin_type=True)


def _ndarray_from_bytes(data: bytes) -> np.ndarray:
    shape, dtype_name, buffer = msgpack.unpackb(data, raw=True)
    name = dtype_name.decode() if isinstance(dtype_name, bytes) else dtype_name
    if name == "bfloat16":
        # flax maps this name to jax.numpy.bfloat16; without ml_dtypes widen to float32 (same values)
        u = np.frombuffer(buffer, dtype=np.uint16).astype(np.uint32) << 16
        return u.view(np.float32).reshape(shape)
    return np.frombuffer(buffer, dtype=np.dtype(name), count=-1, offset=0).reshape(shape, order="C")


def _ext_pack(x):
    if isinstance(x, np.ndarray):
        return msgpack.ExtType(EXT_NDARRAY, _ndarray_to_bytes(x))
    if isinstance(x, np.generic):
        return msgpack.ExtType(EXT_NPSCALAR, _ndarray_to_bytes(np.asarray(x)))
    if isinstance(x, complex):
        return msgpack.ExtType(EXT_NATIVE_COMPLEX, msgpack.packb((x.real, x.imag)))
    return x


def _ext_unpack(code, data):
    if code == EXT_NDARRAY:
        return _ndarray_from_bytes(data)
    if code == EXT_NATIVE_COMPLEX:
        re_, im_ = msgpack.unpackb(data)
        return complex(re_, im_)
    if code == EXT_NPSCALAR:
        return _ndarray_from_bytes(data)[()]
    return msgpack.ExtType(code, data)


def _tuple_to_dict(tpl):
    return {str(i): v for i, v in enumerate(tpl)}


def _chunk(arr: np.ndarray, max_chunk_bytes: int) -> Dict[str, Any]:
    chunksize = max(1, int(max_chunk_bytes / arr.dtype.itemsize))
    flat = arr.reshape(-1)
    chunks = [flat[i : i + chunksize] for i in range(0, flat.size, chunksize)]
    return {_CHUNK_FLAG: True, "shape": _tuple_to_dict(arr.shape), "chunks": _tuple_to_dict(chunks)}


def _chunk_leaves(d, max_chunk_bytes):
    if isinstance(d, dict):
        return {k: _chunk_leaves(v, max_chunk_bytes) for k, v in d.items()}
    if isinstance(d, np.ndarray) and d.size * d.dtype.itemsize > max_chunk_bytes:
        return _chunk(d, max_chunk_bytes)
    return d


def _unchunk_leaves(d):
    if isinstance(d, dict):
        if _CHUNK_FLAG in d:
            shape = tuple(d["shape"][str(i)] for i in range(len(d["shape"])))
            chunks = [d["chunks"][str(i)] for i in range(len(d["chunks"]))]
            return np.concatenate(chunks).reshape(shape)
        return {k: _unchunk_leaves(v) for k, v in d.items()}
    return d


def msgpack_serialize(state_dict: Any, max_chunk_bytes: int = MAX_CHUNK_SIZE) -> bytes:
    """flax.serialization.msgpack_serialize."""
    return msgpack.packb(_chunk_leaves(state_dict, max_chunk_bytes), default=_ext_pack, strict_types=True)


def msgpack_restore(encoded: bytes) -> Any:
    """flax.serialization.msgpack_restore."""
    return _unchunk_leaves(msgpack.unpackb(encoded, ext_hook=_ext_unpack, raw=False, strict_map_key=False))


def to_bytes(target: Any) -> bytes:
    return msgpack_serialize(to_state_dict(target))


# ------------------------------------------------------------------------------------------------ files
_SIGNED_FLOAT_RE = re.compile(r"([-+]?(?:\d+(?:\.\d*)?|\.\d+)(?:[eE][-+]?\d+)?)")


def natural_sort(file_list, signed=True):
    """flax.training.checkpoints.natural_sort: numeric chunks of the name compare as numbers."""
    float_re = _SIGNED_FLOAT_RE if signed else re.compile(r"(\d+(?:\.\d*)?|\.\d+)(?:[eE][-+]?\d+)?")

    def maybe_num(s):
        return float(s) if float_re.fullmatch(s) else s

    def split_keys(s):
        return [maybe_num(c) for c in float_re.split(s)]

    return sorted(file_list, key=split_keys)


def _checkpoint_files(ckpt_dir: str, prefix: str):
    if not os.path.isdir(ckpt_dir):
        return []
    names = [n for n in os.listdir(ckpt_dir) if n.startswith(prefix) and not n.endswith("tmp")]
    return [os.path.join(ckpt_dir, n) for n in natural_sort(names)]


def latest_checkpoint(ckpt_dir: str, prefix: str = "checkpoint_") -> Optional[str]:
    """flax.training.checkpoints.latest_checkpoint (fortuna/training/mixin.py:88-91)."""
    files = _checkpoint_files(str(ckpt_dir), prefix)
    return files[-1] if files else None


def _step_of(path: str, prefix: str) -> float:
    return float(os.path.basename(path)[len(prefix) :])


def save_checkpoint(ckpt_dir: str, target: Any, step, prefix: str = "checkpoint_", keep: int = 1,
                    overwrite: bool = False) -> str:
    """flax.training.checkpoints.save_checkpoint as the reference calls it (fortuna/training/mixin.py:40-47)."""
    ckpt_dir = str(ckpt_dir)
    step = int(step) if float(step) == int(step) else float(step)
    os.makedirs(ckpt_dir, exist_ok=True)
    path = os.path.join(ckpt_dir, f"{prefix}{step}")
    existing = _checkpoint_files(ckpt_dir, prefix)
    if existing and not overwrite and _step_of(existing[-1], prefix) >= float(step):
        raise InvalidCheckpointError(f"Trying to save an outdated checkpoint at step {step}: {existing[-1]} exists")
    tmp = os.path.join(ckpt_dir, f"{prefix}tmp")
    with open(tmp, "wb") as f:
        f.write(to_bytes(target))
    os.replace(tmp, path)
    files = _checkpoint_files(ckpt_dir, prefix)
    if overwrite:  # checkpoints newer than the one just written are stale
        for p in files:
            if _step_of(p, prefix) > float(step):
                os.remove(p)
        files = _checkpoint_files(ckpt_dir, prefix)
    if keep is not None and len(files) > keep:
        for p in files[: len(files) - keep]:
            os.remove(p)
    return path


def restore_checkpoint(ckpt_dir: str, prefix: str = "checkpoint_") -> Optional[Dict]:
    """flax.training.checkpoints.restore_checkpoint(target=None): the raw state dict of the latest checkpoint under a
    directory, or of the given file (fortuna/training/mixin.py:67-83); None when the directory holds none."""
    ckpt_dir = str(ckpt_dir)
    if os.path.isdir(ckpt_dir):
        path = latest_checkpoint(ckpt_dir, prefix)
        if path is None:
            return None
    elif os.path.isfile(ckpt_dir):
        path = ckpt_dir
    else:
        raise ValueError(f"`restore_checkpoint_path={ckpt_dir}` was not found.")
    with open(path, "rb") as f:
        return msgpack_restore(f.read())


def decode_name(d: Dict) -> str:
    """State class name stored as a tuple of code points (fortuna/training/mixin.py:83, utils/strings.py:14-15)."""
    return "".join(chr(int(n)) for n in d["encoded_name"].values())


def encode_name(name: str):
    return tuple(ord(c) for c in name)
