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


def _slot_rows(z, t):
    """What fb200_bma_reduce_cls_shard exports for one shard, restated with the oracle: one row per input,
    [sum_s p (C) | sum_s p^2 (C) | sum_s sum_c p ln p, max_s ll, sum_s exp(ll - max), 0]."""
    p = P.softmax(z)
    logp = P.log_softmax_rows(z)
    ll = P.cls_ensemble_log_prob(z, t)  # [s, B]
    m = ll.max(0)
    rows = np.zeros((z.shape[1], 2 * C + 4), np.float32)
    rows[:, :C] = p.sum(0, dtype=np.float32)
    rows[:, C : 2 * C] = (p * p).sum(0, dtype=np.float32)
    rows[:, 2 * C] = (p * logp).sum((0, 2), dtype=np.float32)
    rows[:, 2 * C + 1] = m
    rows[:, 2 * C + 2] = np.exp(ll - m).sum(0, dtype=np.float32)
    return rows


def _worker(rank, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from fortuna_b200.prob_model.predictive.sharded import SampleShardExchange, local_rows

        z, t = _inputs()
        per = S // WORLD
        rows = _slot_rows(z[rank * per : (rank + 1) * per], t)  # this rank's partials for ALL inputs
        send = torch.from_numpy(rows.reshape(WORLD, B // WORLD, -1).copy())  # slot o = rows owned by rank o
        ex = SampleShardExchange()
        assert (ex.world, ex.rank) == (WORLD, rank)
        recv = ex.exchange_slots(send)
        lo, hi = local_rows(B, WORLD, rank)
        ece = {"counts": torch.tensor([rank + 1, 2], dtype=torch.int32), "correct": torch.tensor([1, 0], dtype=torch.int32),
               "conf_sums": torch.tensor([0.5, 0.25 * (rank + 1)], dtype=torch.float32)}
        ex.allreduce_bins(ece)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, hi=hi, recv=recv.numpy(),
                 counts=ece["counts"].numpy(), conf=ece["conf_sums"].numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_exchange_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    z, t = _inputs()
    want_log_prob = P.cls_log_prob(z, t)
    want_mean = P.cls_mean(z)
    per = S // WORLD
    rows = []
    for rank in range(WORLD):
        d = np.load(tmp_path / f"rank{rank}.npz")
        lo, hi = int(d["lo"]), int(d["hi"])
        rows.append((lo, hi))
        recv = d["recv"]
        assert recv.shape == (WORLD, hi - lo, 2 * C + 4)
        for src in range(WORLD):  # slot `src` = rank src's partials for the rows this rank owns, bit for bit
            np.testing.assert_array_equal(recv[src], _slot_rows(z[src * per : (src + 1) * per], t)[lo:hi])
        # finalise as fb200_bma_finalize_cls_slots does: sums in source order, merged (max, sumexp) pairs
        sp = recv[:, :, :C].sum(0, dtype=np.float32)
        M = recv[:, :, 2 * C + 1].max(0)
        ssum = (recv[:, :, 2 * C + 2] * np.exp(recv[:, :, 2 * C + 1] - M)).sum(0)
        np.testing.assert_allclose(M + np.log(ssum) - np.log(np.float32(S)), want_log_prob[lo:hi], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(sp / np.float32(S), want_mean[lo:hi], rtol=1e-5, atol=1e-7)
        assert d["counts"].tolist() == [3, 4] and np.allclose(d["conf"], [1.0, 0.75])
    assert rows == [(0, B // 2), (B // 2, B)]


def test_local_rows_errors():
    from fortuna_b200.prob_model.predictive.sharded import local_rows

    assert local_rows(12, 4, 3) == (9, 12)
    with pytest.raises(ValueError):
        local_rows(10, 4, 0)
    with pytest.raises(ValueError):
        local_rows(8, 2, 2)


# ------------------------------------------------------------------------------------------------
# Data-parallel sampler step (fb200_sgmcmc_step_dp_f32): host-side split and the semantics it must reproduce
# ------------------------------------------------------------------------------------------------
def test_dp_tile_range_partitions():
    from fortuna_b200 import ops

    for n_tiles in (0, 1, 2, 7, 8, 53441):
        for world in (1, 2, 3, 8):
            r = [ops.dp_tile_range(n_tiles, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n_tiles
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 0


def _dp_worker(rank, port, out_dir):
    """Each rank: own gradient, gloo all-reduce (the reference's lax.pmean), oracle SGHMC step on ITS tile range only,
    then an all-gather of the owned parameter slices - the data movement the fused kernel performs over NVLink."""
    import torch.distributed as dist

    from fortuna_b200 import ops
    from oracle import prng as oprng
    from oracle import sampler as osamp

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    lens = [5000, 7, 2049]
    r = np.random.default_rng(0)
    params = [r.standard_normal(l).astype(np.float32) for l in lens]
    grads = [np.random.default_rng(10 + rank).standard_normal(l).astype(np.float32) for l in lens]
    avg = []
    for g in grads:
        t = torch.from_numpy(g.copy())
        dist.all_reduce(t)
        avg.append((t.numpy() * np.float32(1.0 / WORLD)).astype(np.float32))
    pre = osamp.Preconditioner("identity")
    names = [f"leaf{i}" for i in range(len(lens))]
    state = osamp.sghmc_init(dict(zip(names, params)), oprng.PRNGKey(4), pre)
    upd, _ = osamp.sghmc_update(dict(zip(names, avg)), state, 0.1, osamp.constant_schedule(1e-2), pre)
    new = [p + upd[k] for p, k in zip(params, names)]
    # ownership by tiles of 1024 pairs, as the kernel splits them
    tiles = [(li, ps) for li, l in enumerate(lens) for ps in range(0, (l + 1) // 2, 1024)]
    t0, t1 = ops.dp_tile_range(len(tiles), WORLD, rank)
    mine = [np.zeros(l, bool) for l in lens]
    for li, ps in tiles[t0:t1]:
        h = (lens[li] + 1) // 2
        e = min(ps + 1024, h)
        mine[li][ps:e] = True
        mine[li][h + ps : min(h + e, lens[li])] = True
    out = [np.where(m, n_, np.float32(0)) for m, n_ in zip(mine, new)]
    gathered = []
    for o, m in zip(out, mine):
        t, c = torch.from_numpy(o.copy()), torch.from_numpy(m.astype(np.int32))
        dist.all_reduce(t)
        dist.all_reduce(c)
        assert bool((c == 1).all())  # every element has exactly one owner
        gathered.append(t.numpy())
    np.save(os.path.join(out_dir, f"dp_{rank}.npy"), np.concatenate(gathered))
    if rank == 0:
        np.save(os.path.join(out_dir, "dp_full.npy"), np.concatenate(new))
    dist.destroy_process_group()


def test_dp_step_semantics_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_dp_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    full = np.load(tmp_path / "dp_full.npy")
    for rk in range(WORLD):
        assert np.array_equal(np.load(tmp_path / f"dp_{rk}.npy"), full)


def test_fit_processor_defaults_and_single_process_binding():
    """FitProcessor mirrors fortuna/prob_model/fit_config/processor.py (devices=-1 by default); without an initialised
    multi-rank NCCL group the data-parallel binding is not engaged (checked on the predicate, no GPU needed)."""
    from fortuna_b200.prob_model import FitConfig, FitProcessor

    assert FitConfig().processor.devices == -1 and FitConfig().processor.disable_jit is False
    assert FitConfig(processor=FitProcessor(devices=0)).processor.devices == 0
    assert not dist.is_initialized()


# ------------------------------------------------------------------------------------------------
# Rank-local sampler state of the data-parallel step: who owns what under the two splits, and the gather that makes the
# state whole (phase change of cyclical SGLD with RMSprop, chain checkpoints)
# ------------------------------------------------------------------------------------------------
def _tile_owner_mask(lens, offsets, t0, t1, n):
    """Brute force: elements the kernel touches for tiles [t0, t1) - pair i of a leaf is elements i and i + ceil(len/2)."""
    tiles = [(li, ps) for li, l in enumerate(lens) for ps in range(0, (l + 1) // 2, 1024)]
    m = np.zeros(n, bool)
    for li, ps in tiles[t0:t1]:
        h = (lens[li] + 1) // 2
        e = min(ps + 1024, h)
        m[offsets[li] + ps : offsets[li] + e] = True
        m[offsets[li] + h + ps : offsets[li] + min(h + e, lens[li])] = True
    return m


def test_dp_owned_elem_ranges_match_the_tile_walk():
    from fortuna_b200 import ops
    from fortuna_b200.prob_model.posterior.sgmcmc.dp_step import owned_ranges

    rng = np.random.default_rng(3)
    for lens in ([5000, 7, 2049], [1, 1, 3], [30522 * 8, 768, 2, 3072 * 5 + 1], list(rng.integers(1, 9000, 17))):
        offsets, off = [], 0
        for l in lens:
            offsets.append(off)
            off += (int(l) + 3) // 4 * 4
        n_tiles = sum(((int(l) + 1) // 2 + 1023) // 1024 for l in lens)
        for world in (1, 2, 3, 8):
            seen = np.zeros(off, np.int32)
            for rank in range(world):
                t0, t1 = ops.dp_tile_range(n_tiles, world, rank)
                m = np.zeros(off, bool)
                rs = owned_ranges("tiles", off, world, rank, lens, offsets)
                assert all(a < b for a, b in rs) and all(rs[i][1] < rs[i + 1][0] for i in range(len(rs) - 1))  # merged, sorted
                for a, b in rs:
                    m[a:b] = True
                assert np.array_equal(m, _tile_owner_mask(lens, offsets, t0, t1, off)), (lens[:4], world, rank)
                seen += m
                (e0, e1), = owned_ranges("elems", off, world, rank, lens, offsets)
                assert e0 % 4 == 0 and 0 <= e0 <= e1 <= off
            real = np.zeros(off, bool)
            for l, o in zip(lens, offsets):
                real[o : o + int(l)] = True
            assert np.array_equal(seen == 1, real) and not (seen > 1).any()  # every parameter has exactly one owner


def _gather_worker(rank, port, out_dir):
    import torch.distributed as dist

    from fortuna_b200.prob_model.posterior.sgmcmc.dp_step import gather_owned_ranges, owned_ranges

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    lens, offsets = [5000, 7, 2049], [0, 5000, 5008]
    n = 5008 + 2052
    truth = np.arange(n, dtype=np.float32) + 1
    for scheme in ("tiles", "elems"):
        rs = owned_ranges(scheme, n, WORLD, rank, lens, offsets)
        v = np.full(n, -7.0, np.float32)  # stale values everywhere ...
        for a, b in rs:
            v[a:b] = truth[a:b]           # ... except on the owned pieces
        t = torch.from_numpy(v)
        gather_owned_ranges(t, rs)
        np.save(os.path.join(out_dir, f"g_{scheme}_{rank}.npy"), t.numpy())
    dist.destroy_process_group()


def test_gather_owned_state_world2_gloo(tmp_path):
    """After the gather every rank holds every owner's values (alignment padding, owned by nobody, becomes zero)."""
    port = _free_port()
    mp.spawn(_gather_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    n = 5008 + 2052
    truth = np.arange(n, dtype=np.float32) + 1
    real = np.zeros(n, bool)
    for l, o in zip([5000, 7, 2049], [0, 5000, 5008]):
        real[o : o + l] = True
    for rk in range(WORLD):
        got = np.load(tmp_path / f"g_tiles_{rk}.npy")
        assert np.array_equal(got[real], truth[real]) and not got[~real].any()
        got = np.load(tmp_path / f"g_elems_{rk}.npy")
        assert np.array_equal(got, truth)  # the element split covers the padding too
