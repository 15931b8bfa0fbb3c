"""Checkpoint format (fortuna_b200/training/checkpoint.py): the flax-msgpack wire format restated from flax 0.8.5
(no flax here: the expected bytes below are assembled by hand from the msgpack spec), file naming / retention of
flax.training.checkpoints, and the state-dict shape the reference reads back (fortuna/training/mixin.py:60-85)."""
import os

import numpy as np
import pytest

from fortuna_b200.training import checkpoint as ck


def test_wire_format_of_an_array_leaf_is_flax_ext_type_1():
    arr = np.array([1, 2, 3], dtype=np.int32)
    payload = b"\x93" + b"\x91\x03" + b"\xa5int32" + b"\xc4\x0c" + arr.tobytes()  # (shape, dtype name, bin8 bytes)
    doc = b"\x81" + b"\xa1a" + b"\xc7" + bytes([len(payload)]) + b"\x01" + payload  # {"a": ext8(type 1)}
    assert ck.msgpack_serialize({"a": arr}) == doc
    back = ck.msgpack_restore(doc)
    assert back["a"].dtype == np.int32 and back["a"].tolist() == [1, 2, 3]


def test_wire_format_of_scalars_and_none():
    # python ints / None are native msgpack; a numpy scalar is ext type 3 wrapping a 0-d array
    payload = b"\x93" + b"\x90" + b"\xa7float32" + b"\xc4\x04" + np.float32(1.5).tobytes()
    doc = ck.msgpack_serialize({"n": None, "s": 7, "x": np.float32(1.5)})
    assert len(payload) == 16  # -> msgpack "fixext 16" (0xd8), no length byte
    assert doc == b"\x83" + b"\xa1n\xc0" + b"\xa1s\x07" + b"\xa1x" + b"\xd8\x03" + payload
    back = ck.msgpack_restore(doc)
    assert back["n"] is None and back["s"] == 7 and back["x"] == np.float32(1.5) and back["x"].dtype == np.float32


def test_round_trip_nested_state():
    r = np.random.default_rng(0)
    state = {
        "step": 12,
        "params": {"model": {"params": {"dense": {"kernel": r.standard_normal((3, 4)).astype(np.float32),
                                                  "bias": np.zeros(4, np.float32)}}}},
        "opt_state": {"count": np.asarray(12, np.int32), "rng_key": np.array([1, 2], np.uint32), "momentum": {}},
        "mutable": None,
        "encoded_name": ck.to_state_dict(ck.encode_name("SGHMCState")),
    }
    back = ck.msgpack_restore(ck.msgpack_serialize(state))
    assert back["step"] == 12 and back["mutable"] is None and back["opt_state"]["momentum"] == {}
    np.testing.assert_array_equal(back["params"]["model"]["params"]["dense"]["kernel"],
                                  state["params"]["model"]["params"]["dense"]["kernel"])
    assert back["opt_state"]["rng_key"].dtype == np.uint32 and back["opt_state"]["count"].shape == ()
    assert ck.decode_name(back) == "SGHMCState"  # what fortuna/training/mixin.py:83 does with the dict


def test_chunked_arrays():
    a = np.arange(1000, dtype=np.float32).reshape(10, 100)
    enc = ck.msgpack_serialize({"w": a}, max_chunk_bytes=1024)
    raw = __import__("msgpack").unpackb(enc, ext_hook=ck._ext_unpack, raw=False)
    assert raw["w"]["__msgpack_chunked_array__"] is True and raw["w"]["shape"] == {"0": 10, "1": 100}
    assert len(raw["w"]["chunks"]) == 4 and raw["w"]["chunks"]["0"].shape == (256,)
    np.testing.assert_array_equal(ck.msgpack_restore(enc)["w"], a)


def test_to_state_dict_containers():
    from typing import NamedTuple

    class Inner(NamedTuple):
        pass

    class Outer(NamedTuple):
        count: int
        pre: Inner
        pair: tuple

    d = ck.to_state_dict(Outer(3, Inner(), (np.ones(2), [1, 2])))
    assert d["count"] == 3 and d["pre"] == {} and set(d["pair"]) == {"0", "1"} and d["pair"]["1"] == {"0": 1, "1": 2}


def test_bfloat16_leaves_widen_to_float32():
    import msgpack

    bits = (np.array([1.0, -2.5, 3.140625], np.float32).view(np.uint32) >> 16).astype(np.uint16)
    ext = msgpack.ExtType(1, msgpack.packb(((3,), "bfloat16", bits.tobytes()), use_bin_type=True))
    got = ck.msgpack_restore(msgpack.packb({"w": ext}))
    assert got["w"].dtype == np.float32 and got["w"].tolist() == [1.0, -2.5, 3.140625]


def test_save_restore_latest_keep_overwrite(tmp_path):
    d = str(tmp_path / "c")
    for step in (1, 2, 9, 10):
        ck.save_checkpoint(d, {"step": step, "x": np.full(2, step, np.float32)}, step=step, keep=2)
    assert sorted(os.listdir(d)) == ["checkpoint_10", "checkpoint_9"]  # natural order: 10 after 9
    assert ck.latest_checkpoint(d).endswith("checkpoint_10")
    assert ck.restore_checkpoint(d)["step"] == 10
    assert ck.restore_checkpoint(os.path.join(d, "checkpoint_9"))["step"] == 9
    with pytest.raises(ck.InvalidCheckpointError):
        ck.save_checkpoint(d, {"step": 5}, step=5, keep=2)
    ck.save_checkpoint(d, {"step": 5}, step=5, keep=1, overwrite=True)  # force_save: newer ones are dropped
    assert os.listdir(d) == ["checkpoint_5"]
    os.makedirs(tmp_path / "empty")
    assert ck.restore_checkpoint(str(tmp_path / "empty")) is None
    with pytest.raises(ValueError):
        ck.restore_checkpoint(str(tmp_path / "missing"))
    assert ck.latest_checkpoint(str(tmp_path / "missing")) is None


def test_posterior_state_dict_matches_reference_fields():
    """Field set of SGHMCState's to_state_dict (posterior/state.py:24-36, sghmc_state.py:23-32, flax TrainState) and the
    sub-layouts of OptaxSGHMCState / OptaxCyclicalSGLDState."""
    import torch

    from fortuna_b200.prob_model.posterior.sgmcmc.cyclical_sgld.cyclical_sgld_integrator import OptaxCyclicalSGLDState
    from fortuna_b200.prob_model.posterior.sgmcmc.sghmc.sghmc_integrator import OptaxSGHMCState
    from fortuna_b200.prob_model.posterior.sgmcmc.sgmcmc_posterior_state_repository import posterior_state_dict
    from fortuna_b200.prob_model.posterior.sgmcmc.sgmcmc_preconditioner import (
        IdentityPreconditionerState,
        RMSPropPreconditionerState,
    )
    from fortuna_b200.training.train_state import TrainState
    from fortuna_b200.utils.flat_tree import FlatTree

    tree = {"model": {"params": {"output_subnet": {"last": {"kernel": np.ones((3, 2), np.float32), "bias": np.zeros(2, np.float32)}}}}}
    p = FlatTree.from_tree(tree, device="cpu")
    key = torch.tensor([-1, 5], dtype=torch.int32)
    inner = OptaxSGHMCState(count=4, rng_key=key, momentum=p.zeros_like(),
                            preconditioner_state=RMSPropPreconditionerState(grad_moment_estimates=p.zeros_like()))
    st = TrainState(step=4, params=p, tx=None, opt_state=inner)
    which = (["model", "params", "output_subnet", "last", "bias"], ["model", "params", "output_subnet", "last", "kernel"])
    d = ck.msgpack_restore(ck.msgpack_serialize(posterior_state_dict(st, "SGHMCState", which)))
    assert set(d) == {"step", "params", "opt_state", "encoded_name", "frozen_params", "dynamic_scale", "mutable",
                      "calib_params", "calib_mutable", "grad_accumulated", "_encoded_which_params"}
    assert d["step"] == 4 and ck.decode_name(d) == "SGHMCState"
    assert d["params"]["model"]["params"]["output_subnet"]["last"]["kernel"].shape == (3, 2)
    o = d["opt_state"]
    assert set(o) == {"count", "rng_key", "momentum", "preconditioner_state"}
    assert o["count"].dtype == np.int32 and int(o["count"]) == 4
    assert o["rng_key"].dtype == np.uint32 and o["rng_key"].tolist() == [0xFFFFFFFF, 5]
    assert set(o["preconditioner_state"]) == {"grad_moment_estimates"}
    # utils/strings.py:29-43 decoding of the stored key paths
    dec = tuple(["".join(chr(int(c)) for c in v) for v in path.values()] for path in d["_encoded_which_params"].values())
    assert dec == tuple(which)
    cyc = TrainState(step=1, params=p, tx=None, opt_state=OptaxCyclicalSGLDState(
        sgd_state=((), ()), sgld_state=inner._replace(preconditioner_state=IdentityPreconditionerState())))
    d2 = posterior_state_dict(cyc, "CyclicalSGLDState")
    assert d2["opt_state"]["sgd_state"] == {"0": {}, "1": {}} and d2["opt_state"]["sgld_state"]["preconditioner_state"] == {}
    assert d2["_encoded_which_params"] is None
