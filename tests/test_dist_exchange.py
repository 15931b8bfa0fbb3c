"""World-size-2 gloo test (CPU) of the S-sharded predictive's exchange step
(fortuna_b200/prob_model/predictive/sharded.py): the collective plumbing - which rows each rank ends up with,
in which order the gathered (max, sumexp) pairs arrive, that the summed partials are the sums over ranks - is
checked against the oracle evaluated on ALL samples.  The per-rank partial sums fed in are formed from the
oracle (the CUDA kernels that produce them on a GPU box are covered by tests/test_gpu_predictive.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import predictive as P

S, B, C, WORLD = 6, 16, 5, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs():
    r = np.random.default_rng(42)
    z = (3 * r.standard_normal((S, B, C))).astype(np.float32)
    t = r.integers(0, C, B).astype(np.int32)
    return z, t


def _partials(z, t):
    """What fb200_bma_reduce_cls (partial mode) exports for one shard, restated with the oracle."""
    p = P.softmax(z)
    logp = P.log_softmax_rows(z)
    ll = P.cls_ensemble_log_prob(z, t)  # [s, B]
    m = ll.max(0)
    return {
        "sum_p": p.sum(0, dtype=np.float32),
        "sum_p2": (p * p).sum(0, dtype=np.float32),
        "sum_pq": (p * (1 - p)).sum(0, dtype=np.float32),
        "sum_plogp": (p * logp).sum((0, 2), dtype=np.float32),
        "ll_max": m.astype(np.float32),
        "ll_sumexp": np.exp(ll - m).sum(0, dtype=np.float32),
    }


def _worker(rank, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from fortuna_b200.prob_model.predictive.sharded import SampleShardExchange, local_rows

        z, t = _inputs()
        per = S // WORLD
        shard = z[rank * per : (rank + 1) * per]
        part = {k: torch.from_numpy(v) for k, v in _partials(shard, t).items()}
        ex = SampleShardExchange()
        assert (ex.world, ex.rank) == (WORLD, rank)
        summed, gmax, gsum = ex.exchange_partials(part)
        lo, hi = local_rows(B, WORLD, rank)
        ece = {"counts": torch.tensor([rank + 1, 2], dtype=torch.int32), "correct": torch.tensor([1, 0], dtype=torch.int32),
               "conf_sums": torch.tensor([0.5, 0.25 * (rank + 1)], dtype=torch.float32)}
        ex.allreduce_bins(ece)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, hi=hi, gmax=gmax.numpy(), gsum=gsum.numpy(),
                 counts=ece["counts"].numpy(), conf=ece["conf_sums"].numpy(),
                 **{k: v.numpy() for k, v in summed.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_exchange_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    z, t = _inputs()
    full = _partials(z, t)
    want_log_prob = P.cls_log_prob(z, t)
    want_mean = P.cls_mean(z)
    rows = []
    for rank in range(WORLD):
        d = np.load(tmp_path / f"rank{rank}.npz")
        lo, hi = int(d["lo"]), int(d["hi"])
        rows.append((lo, hi))
        for k in ("sum_p", "sum_p2", "sum_pq", "sum_plogp"):
            np.testing.assert_allclose(d[k], full[k][lo:hi], rtol=2e-6, atol=1e-6, err_msg=k)
        # rank-major gather, sliced to the local rows
        assert d["gmax"].shape == (WORLD, hi - lo)
        per = S // WORLD
        for r in range(WORLD):
            pr = _partials(z[r * per : (r + 1) * per], t)
            np.testing.assert_array_equal(d["gmax"][r], pr["ll_max"][lo:hi])
            np.testing.assert_array_equal(d["gsum"][r], pr["ll_sumexp"][lo:hi])
        # merge + finalise as the kernels do: M = max_r m_r, sum_r s_r exp(m_r - M); log_prob = M + log sum - log S
        M = d["gmax"].max(0)
        ssum = (d["gsum"] * np.exp(d["gmax"] - M)).sum(0)
        np.testing.assert_allclose(M + np.log(ssum) - np.log(np.float32(S)), want_log_prob[lo:hi], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(d["sum_p"] / np.float32(S), want_mean[lo:hi], rtol=1e-5, atol=1e-7)
        assert d["counts"].tolist() == [3, 4] and np.allclose(d["conf"], [1.0, 0.75])
    assert rows == [(0, B // 2), (B // 2, B)]


def test_local_rows_errors():
    from fortuna_b200.prob_model.predictive.sharded import local_rows

    assert local_rows(12, 4, 3) == (9, 12)
    with pytest.raises(ValueError):
        local_rows(10, 4, 0)
    with pytest.raises(ValueError):
        local_rows(8, 2, 2)
