"""Torch-pointer driver over the C ABI: every function here is one library call on the current CUDA
stream with caller-visible torch tensors as storage.  No arithmetic happens in Python or in torch."""

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import BF16, F32, W_SCD, W_SDC, check, lib

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, dtype=None, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise ValueError("a required tensor is None")
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError("fortuna_b200 operates on CUDA tensors only (there is no CPU path)")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    if dtype is not None and t.dtype not in (dtype if isinstance(dtype, tuple) else (dtype,)):
        raise ValueError(f"expected dtype {dtype}, got {t.dtype}")
    return C.c_void_p(t.data_ptr())


KEY_DTYPE = torch.int32  # uint32[2] bit patterns (torch has no general uint32 arithmetic; none is needed)


def key_from_host(words, device):
    """uint32[2] host words -> device key tensor."""
    a = np.asarray(words, dtype=np.uint32).view(np.int32)
    return torch.from_numpy(a.copy()).to(device)


def key_to_host(key):
    return key.detach().cpu().numpy().view(np.uint32).copy()


def PRNGKey(seed, device="cuda"):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return key_from_host([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], device)


# ------------------------------------------------------------------------------------------- PRNG
def split(key, num=2):
    out = torch.empty((num, 2), dtype=KEY_DTYPE, device=key.device)
    check(lib().fb200_threefry_split(_ptr(key, KEY_DTYPE), num, _ptr(out), _stream()), "fb200_threefry_split")
    return out


def fold_in(key, data):
    out = torch.empty(2, dtype=KEY_DTYPE, device=key.device)
    check(
        lib().fb200_threefry_fold_in(_ptr(key, KEY_DTYPE), int(data) & 0xFFFFFFFF, _ptr(out), _stream()),
        "fb200_threefry_fold_in",
    )
    return out


def random_bits(key, n):
    out = torch.empty(n, dtype=torch.int32, device=key.device)
    check(lib().fb200_random_bits(_ptr(key, KEY_DTYPE), n, _ptr(out), _stream()), "fb200_random_bits")
    return out


def random_normal(key, n):
    out = torch.empty(n, dtype=torch.float32, device=key.device)
    check(lib().fb200_random_normal(_ptr(key, KEY_DTYPE), n, _ptr(out), _stream()), "fb200_random_normal")
    return out


def sample_indices(key, n_draws, n_samples):
    out = torch.empty(n_draws, dtype=torch.int32, device=key.device)
    check(
        lib().fb200_sample_indices(_ptr(key, KEY_DTYPE), n_draws, n_samples, _ptr(out), _stream()),
        "fb200_sample_indices",
    )
    return out


def sample_indices_from_keys(keys, n_samples):
    """keys int32[n,2] -> int32[n]: choice(keys[i], n_samples) (no split over draws)."""
    n = keys.shape[0]
    out = torch.empty(n, dtype=torch.int32, device=keys.device)
    check(
        lib().fb200_choice_from_keys(_ptr(keys, KEY_DTYPE), n, n_samples, _ptr(out), _stream()),
        "fb200_choice_from_keys",
    )
    return out


# ------------------------------------------------------------------------------------------- sampler
def build_leaf_tables(lens, offsets, device):
    """Host leaf/tile tables -> device int tensors (leaves [L,2] int64, tiles [T,2] int32)."""
    n_leaves = len(lens)
    leaves = (_lib.Leaf * n_leaves)()
    for i, (o, l) in enumerate(zip(offsets, lens)):
        leaves[i].offset = int(o)
        leaves[i].len = int(l)
    n_tiles = lib().fb200_sgmcmc_num_tiles_host(leaves, n_leaves)
    if n_tiles < 0:
        check(int(n_tiles), "fb200_sgmcmc_num_tiles_host")
    tiles = (_lib.Tile * n_tiles)()
    check(lib().fb200_sgmcmc_build_tiles_host(leaves, n_leaves, tiles, n_tiles), "fb200_sgmcmc_build_tiles_host")
    leaves_np = np.frombuffer(leaves, dtype=np.int64).reshape(n_leaves, 2).copy()
    tiles_np = np.frombuffer(tiles, dtype=np.int32).reshape(n_tiles, 2).copy()
    return torch.from_numpy(leaves_np).to(device), torch.from_numpy(tiles_np).to(device)


def grad_clip_mult(grad, max_grad_norm, params=None, joint=None, ws_cache=None):
    """float32 [1] device tensor: min(1, max_grad_norm / (1e-7 + ||g||)) - clip_grandients_by_norm's multiplier
    (fortuna/utils/training.py:14-20) for the gradient the step is about to use.  ``joint = (lik_scale, prior_prec)``: g is
    the log-joint gradient the fused mode forms from ``grad`` and ``params``.  Pass the result as
    ``make_hparams(..., grad_mult=...)``: the step kernel scales the gradient as it loads it."""
    n = grad.numel()
    nws = int(lib().fb200_grad_clip_workspace_doubles(n))
    ws = ws_cache.get("clip_ws") if ws_cache is not None else None
    if ws is None or ws.numel() < nws or ws.device != grad.device:
        ws = torch.empty(nws, dtype=torch.float64, device=grad.device)
        if ws_cache is not None:
            ws_cache["clip_ws"] = ws
    mult = torch.empty(1, dtype=torch.float32, device=grad.device)
    jg = _joint_struct(joint)
    check(
        lib().fb200_grad_clip_mult(_ptr(grad, (torch.float32, torch.bfloat16)), _DT[grad.dtype], _ptr(params, torch.float32, True),
                                   C.byref(jg) if jg is not None else None, n, float(np.float32(max_grad_norm)),
                                   _ptr(ws, torch.float64), _ptr(mult), _stream()),
        "fb200_grad_clip_mult",
    )
    return mult


def make_hparams(step_size, momentum_decay=0.0, rmsprop=False, rms_alpha=0.99, rms_eps=1e-7, zero_momentum=False,
                 grad_mult=None, grad_dtype=torch.float32):
    """``grad_mult``: optional device float32[1] (see ``grad_clip_mult``); the caller keeps it alive across the step."""
    hp = _lib.SgmcmcHParams()
    hp.grad_mult = grad_mult.data_ptr() if grad_mult is not None else None
    hp.grad_dtype = _DT[grad_dtype]
    hp.step_size = float(np.float32(step_size))
    hp.momentum_decay = float(momentum_decay)
    hp.zero_momentum = int(bool(zero_momentum))
    hp.rmsprop = int(bool(rmsprop))
    hp.rms_alpha = float(rms_alpha)
    hp.rms_eps = float(rms_eps)
    return hp


def _joint_args(joint, n_ws, device, ws_cache):
    """joint = (lik_scale, prior_prec, theta_sq_out or None) -> (struct, workspace ptr, out ptr)."""
    lik_scale, prior_prec, sq_out = joint
    jg = _lib.JointGrad()
    jg.lik_scale = float(np.float32(lik_scale))
    jg.prior_prec = float(np.float32(prior_prec))
    ws = None
    if sq_out is not None:
        ws = ws_cache.get("sq_ws") if ws_cache is not None else None
        if ws is None or ws.numel() < n_ws or ws.device != device:
            ws = torch.empty(n_ws, dtype=torch.float64, device=device)
            if ws_cache is not None:
                ws_cache["sq_ws"] = ws
    return jg, _ptr(ws, torch.float64, True), _ptr(sq_out, torch.float64, True)


def sgmcmc_step(params, grad, momentum, precond_v, updates_out, leaves, tiles, key_in, key_out, leaf_keys_ws, hp,
                joint=None, ws_cache=None):
    """joint = (n_data / batch_size, prior precision, float64[1] tensor or None): `grad` is the gradient of the batch
    log-likelihood sum and the kernel forms the log-joint gradient itself (fb200_sgmcmc_step_joint_f32)."""
    n = grad.numel()
    if joint is not None:
        if updates_out is not None or params is None:
            raise ValueError("the fused log-joint mode updates params in place (no `updates` output)")
        jg, wsp, outp = _joint_args(joint, tiles.shape[0], grad.device, ws_cache)
        check(
            lib().fb200_sgmcmc_step_joint_f32(
                _ptr(params, torch.float32), _ptr(grad, (torch.float32, torch.bfloat16)), _ptr(momentum, torch.float32),
                _ptr(precond_v, torch.float32, True), n, _ptr(leaves, torch.int64), leaves.shape[0],
                _ptr(tiles, torch.int32), tiles.shape[0], _ptr(key_in, KEY_DTYPE), _ptr(key_out, KEY_DTYPE),
                _ptr(leaf_keys_ws, KEY_DTYPE), C.byref(hp), C.byref(jg), wsp, outp, _stream(),
            ),
            "fb200_sgmcmc_step_joint_f32",
        )
        return
    check(
        lib().fb200_sgmcmc_step_f32(
            _ptr(params, torch.float32, True),
            _ptr(grad, (torch.float32, torch.bfloat16)),
            _ptr(momentum, torch.float32),
            _ptr(precond_v, torch.float32, True),
            _ptr(updates_out, torch.float32, True),
            n,
            _ptr(leaves, torch.int64),
            leaves.shape[0],
            _ptr(tiles, torch.int32),
            tiles.shape[0],
            _ptr(key_in, KEY_DTYPE),
            _ptr(key_out, KEY_DTYPE),
            _ptr(leaf_keys_ws, KEY_DTYPE),
            C.byref(hp),
            _stream(),
        ),
        "fb200_sgmcmc_step_f32",
    )


def dp_tile_range(n_tiles: int, world: int, rank: int):
    """Contiguous, balanced split of the tile table: rank r owns tiles [begin, end)."""
    base, rem = divmod(int(n_tiles), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def dp_elem_range(n: int, world: int, rank: int):
    """Contiguous split of [0, n) into 4-aligned element ranges: rank r owns [begin, end)."""
    per = ((int(n) + 3) // 4 + world - 1) // world * 4
    return min(rank * per, int(n)), min((rank + 1) * per, int(n))


def dp_owned_elem_ranges(lens, offsets, tile_range, tile_pairs=_lib.TILE_PAIRS):
    """Flat element ranges [(begin, end), ...] a rank owns under the TILE split of the data-parallel step: tile t of a
    leaf of length n covers the counter pairs [1024 t, 1024 (t + 1)) of that leaf, i.e. the elements i and i + ceil(n/2)
    of those pairs (jax's threefry layout, see csrc/sgmcmc.cu) - so a contiguous range of tiles is whole leaves plus, at
    its two ends, a lower-half and an upper-half piece of a partially owned leaf.  Adjacent ranges are merged."""
    t0, t1 = tile_range
    out, base = [], 0
    for n, off in zip(lens, offsets):
        h = (int(n) + 1) // 2
        nt = (h + tile_pairs - 1) // tile_pairs
        a, b = max(t0, base), min(t1, base + nt)
        if a < b:
            p0, p1 = (a - base) * tile_pairs, min((b - base) * tile_pairs, h)
            for lo, hi in ((off + p0, off + p1), (off + h + p0, off + min(h + p1, int(n)))):
                if hi > lo:
                    if out and out[-1][1] == lo:
                        out[-1] = (out[-1][0], hi)
                    else:
                        out.append((lo, hi))
        base += nt
    return out


def _dp_peers(grad_ptrs, param_ptrs, rank, mc_grad_ptr=0, mc_param_ptr=0):
    world = len(grad_ptrs)
    peers = _lib.DpPeers()
    peers.n_ranks, peers.rank = world, int(rank)
    for p in range(world):
        peers.grad[p] = int(grad_ptrs[p])
        peers.params[p] = int(param_ptrs[p])
    if mc_grad_ptr and mc_param_ptr:  # NVLS multicast addresses: in-switch gradient sum and parameter replication
        peers.mc_grad, peers.mc_params = int(mc_grad_ptr), int(mc_param_ptr)
    return peers


def _joint_struct(joint):
    if joint is None:
        return None
    jg = _lib.JointGrad()
    jg.lik_scale = float(np.float32(joint[0]))
    jg.prior_prec = float(np.float32(joint[1]))
    return jg


def sgd_explore_step_dp(grad_ptrs, param_ptrs, rank, precond_v, n, hp, joint=None, elem_range=None, mc_grad_ptr=0,
                        mc_param_ptr=0):
    """Data-parallel exploration step of cyclical SGLD (fb200_sgd_explore_step_dp_f32); see ``sgmcmc_step_dp``."""
    peers = _dp_peers(grad_ptrs, param_ptrs, rank, mc_grad_ptr, mc_param_ptr)
    e0, e1 = elem_range if elem_range is not None else dp_elem_range(n, len(grad_ptrs), rank)
    jg = _joint_struct(joint)
    check(
        lib().fb200_sgd_explore_step_dp_f32(_ptr(precond_v, torch.float32, True), int(n), int(e0), int(e1), C.byref(hp),
                                            C.byref(jg) if jg is not None else None, C.byref(peers), _stream()),
        "fb200_sgd_explore_step_dp_f32",
    )


def sgmcmc_step_dp(grad_ptrs, param_ptrs, rank, momentum, precond_v, n, leaves, tiles, key_in, key_out, leaf_keys_ws,
                   hp, joint=None, tile_range=None, mc_grad_ptr=0, mc_param_ptr=0):
    """Data-parallel fused step (fb200_sgmcmc_step_dp_f32): ``grad_ptrs`` / ``param_ptrs`` are every rank's [n] float32
    buffers as peer-mapped device addresses (rank order); this rank updates tiles ``tile_range`` (default: its balanced
    share) reading all gradients and writing all parameter replicas.  The caller brackets the call with barriers."""
    world = len(grad_ptrs)
    peers = _dp_peers(grad_ptrs, param_ptrs, rank, mc_grad_ptr, mc_param_ptr)
    t0, t1 = tile_range if tile_range is not None else dp_tile_range(tiles.shape[0], world, rank)
    jg = _joint_struct(joint)
    check(
        lib().fb200_sgmcmc_step_dp_f32(
            _ptr(momentum, torch.float32), _ptr(precond_v, torch.float32, True), int(n), _ptr(leaves, torch.int64),
            leaves.shape[0], _ptr(tiles, torch.int32), int(t0), int(t1), _ptr(key_in, KEY_DTYPE), _ptr(key_out, KEY_DTYPE),
            _ptr(leaf_keys_ws, KEY_DTYPE), C.byref(hp), C.byref(jg) if jg is not None else None, C.byref(peers), _stream(),
        ),
        "fb200_sgmcmc_step_dp_f32",
    )


def sgd_explore_step(params, grad, precond_v, updates_out, hp, joint=None, ws_cache=None):
    if joint is not None:
        if updates_out is not None or params is None:
            raise ValueError("the fused log-joint mode updates params in place (no `updates` output)")
        n_ws = int(lib().fb200_sgd_explore_joint_workspace_doubles(grad.numel()))
        jg, wsp, outp = _joint_args(joint, n_ws, grad.device, ws_cache)
        check(
            lib().fb200_sgd_explore_step_joint_f32(
                _ptr(params, torch.float32), _ptr(grad, torch.float32), _ptr(precond_v, torch.float32, True),
                grad.numel(), C.byref(hp), C.byref(jg), wsp, outp, _stream(),
            ),
            "fb200_sgd_explore_step_joint_f32",
        )
        return
    check(
        lib().fb200_sgd_explore_step_f32(
            _ptr(params, torch.float32, True),
            _ptr(grad, torch.float32),
            _ptr(precond_v, torch.float32, True),
            _ptr(updates_out, torch.float32, True),
            grad.numel(),
            C.byref(hp),
            _stream(),
        ),
        "fb200_sgd_explore_step_f32",
    )


def bank_put(bank, row, src):
    check(
        lib().fb200_bank_put_f32(_ptr(bank, torch.float32), bank.shape[1], int(row), _ptr(src, torch.float32), _stream()),
        "fb200_bank_put_f32",
    )


# ------------------------------------------------------------------------------------------- last layer
def last_layer_logits(x, w, bias=None, log_temp=None, out_dtype=torch.float32, w_layout=W_SDC, path=0, out=None,
                      sample_idx=None):
    """x [B,D], w [n,D,C] (W_SDC) or [n,C,D] (W_SCD), bias [n,C] f32 -> out [S,B,C].
    sample_idx (int32 [S], device) selects which of the n stored samples each output row uses; None = all n."""
    if x.dtype != w.dtype:
        raise ValueError("x and w must share a dtype")
    S = w.shape[0] if sample_idx is None else sample_idx.numel()
    B, D = x.shape
    Cn = w.shape[2] if w_layout == W_SDC else w.shape[1]
    if (w.shape[1] if w_layout == W_SDC else w.shape[2]) != D:
        raise ValueError("feature dimension mismatch between x and w")
    if out is None:
        out = torch.empty((S, B, Cn), dtype=out_dtype, device=x.device)
    check(
        lib().fb200_last_layer_logits(
            _ptr(x, (torch.float32, torch.bfloat16)),
            _ptr(w),
            _ptr(bias, torch.float32, True),
            _ptr(log_temp, torch.float32, True),
            _ptr(sample_idx, torch.int32, True),
            w.shape[0],
            _ptr(out, (torch.float32, torch.bfloat16)),
            S,
            B,
            D,
            Cn,
            _DT[x.dtype],
            _DT[out.dtype],
            w_layout,
            path,
            _stream(),
        ),
        "fb200_last_layer_logits",
    )
    return out


EPI_LOGITS, EPI_SOFTMAX, EPI_LOG_SOFTMAX = 0, 1, 2


def last_layer_logits_ex(x, w, bias=None, log_temp=None, out_dtype=torch.float32, w_layout=W_SDC, epilogue=EPI_LOGITS,
                         want_lse=False, sample_idx=None):
    """The last-layer contraction with the softmax / log-softmax epilogue and / or the per-row logsumexp over classes
    (fb200_last_layer_logits_ex).  Returns out [S,B,C], or (out, lse [S,B]) with want_lse."""
    if x.dtype != w.dtype:
        raise ValueError("x and w must share a dtype")
    S = w.shape[0] if sample_idx is None else sample_idx.numel()
    B, D = x.shape
    Cn = w.shape[2] if w_layout == W_SDC else w.shape[1]
    if (w.shape[1] if w_layout == W_SDC else w.shape[2]) != D:
        raise ValueError("feature dimension mismatch between x and w")
    out = torch.empty((S, B, Cn), dtype=out_dtype, device=x.device)
    lse = torch.empty((S, B), dtype=torch.float32, device=x.device) if want_lse else None
    nbytes = int(lib().fb200_last_layer_ex_workspace_bytes(S, B, Cn)) if want_lse else 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device) if nbytes else None
    check(
        lib().fb200_last_layer_logits_ex(
            _ptr(x, (torch.float32, torch.bfloat16)), _ptr(w), _ptr(bias, torch.float32, True),
            _ptr(log_temp, torch.float32, True), _ptr(sample_idx, torch.int32, True), w.shape[0],
            _ptr(out, (torch.float32, torch.bfloat16)), S, B, D, Cn, _DT[x.dtype], _DT[out.dtype], w_layout,
            int(epilogue), _ptr(lse, torch.float32, True), _ptr(ws, None, True), nbytes, _stream(),
        ),
        "fb200_last_layer_logits_ex",
    )
    return (out, lse) if want_lse else out


def log_prob_from_lse(logits, lse, targets, want_ensemble=False):
    """log_prob [B] (and ensemble_log_prob [S,B]) from calibrated logits [S,B,C] and their row logsumexp [S,B]."""
    S, B, Cn = logits.shape
    lp = torch.empty(B, dtype=torch.float32, device=logits.device)
    ens = torch.empty((S, B), dtype=torch.float32, device=logits.device) if want_ensemble else None
    check(
        lib().fb200_log_prob_from_lse(
            _ptr(logits, (torch.float32, torch.bfloat16)), _DT[logits.dtype], _ptr(lse, torch.float32),
            _ptr(targets, torch.int32), S, B, Cn, _ptr(ens, torch.float32, True), _ptr(lp), _stream(),
        ),
        "fb200_log_prob_from_lse",
    )
    return (lp, ens) if want_ensemble else lp


def transpose_w_bf16(w):
    S, D, Cn = w.shape
    wt = torch.empty((S, Cn, D), dtype=torch.bfloat16, device=w.device)
    check(
        lib().fb200_transpose_w_bf16(_ptr(w, (torch.float32, torch.bfloat16)), _DT[w.dtype], _ptr(wt), S, D, Cn, _stream()),
        "fb200_transpose_w_bf16",
    )
    return wt


# ------------------------------------------------------------------------------------------- reductions
CLS_FIELDS = tuple(n for n, _ in _lib.ClsOutputs._fields_)
_CLS_BC = ("mean", "aleatoric_variance", "epistemic_variance", "variance", "std")
_CLS_B = ("entropy", "aleatoric_entropy", "epistemic_entropy", "log_prob", "confidence", "brier_terms")


def _alloc_cls_outputs(want, S, B, Cn, device):
    outs = {}
    for name in want:
        if name in _CLS_BC:
            outs[name] = torch.empty((B, Cn), dtype=torch.float32, device=device)
        elif name in _CLS_B:
            outs[name] = torch.empty(B, dtype=torch.float32, device=device)
        elif name == "mode":
            outs[name] = torch.empty(B, dtype=torch.int32, device=device)
        elif name == "ensemble_log_prob":
            outs[name] = torch.empty((S, B), dtype=torch.float32, device=device)
        else:
            raise ValueError(f"unknown classification output {name!r}")
    return outs


def _cls_struct(outs):
    st = _lib.ClsOutputs()
    for name, t in outs.items():
        setattr(st, name, t.data_ptr())
    return st


def bma_reduce_cls(logits, want, targets=None, log_temp=None, out=None):
    """One pass over logits [S,B,C] (f32 or bf16). `want`: iterable of output names -> dict of tensors."""
    S, B, Cn = logits.shape
    outs = out if out is not None else _alloc_cls_outputs(tuple(want), S, B, Cn, logits.device)
    st = _cls_struct(outs)
    if Cn > 1024:  # rows wider than a warp's registers: the two-pass kernel
        nbytes = int(lib().fb200_bma_reduce_cls_wide_workspace_bytes(S, B))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=logits.device)
        check(
            lib().fb200_bma_reduce_cls_wide(
                _ptr(logits, (torch.float32, torch.bfloat16)), _DT[logits.dtype], _ptr(log_temp, torch.float32, True),
                _ptr(targets, torch.int32, True), S, B, Cn, C.byref(st), _ptr(ws), nbytes, _stream(),
            ),
            "fb200_bma_reduce_cls_wide",
        )
        return outs
    check(
        lib().fb200_bma_reduce_cls(
            _ptr(logits, (torch.float32, torch.bfloat16)),
            _DT[logits.dtype],
            _ptr(log_temp, torch.float32, True),
            _ptr(targets, torch.int32, True),
            S,
            B,
            Cn,
            C.byref(st),
            _stream(),
        ),
        "fb200_bma_reduce_cls",
    )
    return outs


def cls_slot_row_floats(Cn):
    """Floats per row of a shard slot: [sum p (C) | sum p^2 (C) | sum plogp, ll_max, ll_sumexp, 0]."""
    return int(lib().fb200_cls_slot_row_floats(int(Cn)))


def alloc_cls_slots(n_ranks, rows_per_rank, Cn, device):
    """[n_ranks, rows_per_rank, row_floats] float32: a send buffer (slot o -> rank o) or a receive buffer."""
    return torch.empty((n_ranks, rows_per_rank, cls_slot_row_floats(Cn)), dtype=torch.float32, device=device)


def slot_pointer_table(slot_tensors_or_ptrs, device):
    """Device array of float* (uint64) from tensors (local memory) or raw integer addresses (peer memory)."""
    ptrs = [int(t.data_ptr()) if isinstance(t, torch.Tensor) else int(t) for t in slot_tensors_or_ptrs]
    return torch.from_numpy(np.array(ptrs, dtype=np.uint64).view(np.int64).copy()).to(device)


def bma_reduce_cls_shard(logits, slot_ptrs, targets=None, log_temp=None, rank=0, ensemble_log_prob=None):
    """Shard mode of the single-pass reduction: export this rank's partial sums row by row into the slot of the
    rank owning the row.  slot_ptrs: int64 device tensor [n_ranks] from slot_pointer_table."""
    S, B, Cn = logits.shape
    check(
        lib().fb200_bma_reduce_cls_shard(
            _ptr(logits, (torch.float32, torch.bfloat16)),
            _DT[logits.dtype],
            _ptr(log_temp, torch.float32, True),
            _ptr(targets, torch.int32, True),
            S,
            B,
            Cn,
            slot_ptrs.numel(),
            int(rank),
            _ptr(slot_ptrs, torch.int64),
            _ptr(ensemble_log_prob, torch.float32, True),
            _stream(),
        ),
        "fb200_bma_reduce_cls_shard",
    )


def bma_finalize_cls_slots(slots, s_total, Cn, want, out=None, targets=None):
    """slots [n_src, rows, row_floats] (received, one per source rank) -> final outputs for the owned rows; `targets`
    (int32 [rows], the owned rows') only for ``brier_terms``."""
    n_src, rows, _ = slots.shape
    outs = out if out is not None else _alloc_cls_outputs(tuple(want), s_total, rows, Cn, slots.device)
    ost = _cls_struct(outs)
    check(
        lib().fb200_bma_finalize_cls_slots(_ptr(slots, torch.float32), n_src, rows, Cn, int(s_total), C.byref(ost),
                                           _ptr(targets, torch.int32, True), _stream()),
        "fb200_bma_finalize_cls_slots",
    )
    return outs


def variance_combine(aleatoric, epistemic=None, want_std=False):
    """variance = aleatoric + epistemic (or `aleatoric` itself when epistemic is None), optionally its square root."""
    out = torch.empty_like(aleatoric)
    check(
        lib().fb200_variance_combine(_ptr(aleatoric, torch.float32), _ptr(epistemic, torch.float32, True), aleatoric.numel(),
                                     None if want_std else _ptr(out), _ptr(out) if want_std else None, _stream()),
        "fb200_variance_combine",
    )
    return out


def bma_reduce_reg(outputs, want, targets=None, log_temp=None):
    S, B, two_d = outputs.shape
    d = two_d // 2
    outs = {}
    for name in want:
        if name in ("mean", "aleatoric_variance", "epistemic_variance", "variance", "std"):
            outs[name] = torch.empty((B, d), dtype=torch.float32, device=outputs.device)
        elif name == "log_prob":
            outs[name] = torch.empty(B, dtype=torch.float32, device=outputs.device)
        elif name == "ensemble_log_prob":
            outs[name] = torch.empty((S, B), dtype=torch.float32, device=outputs.device)
        else:
            raise ValueError(f"unknown regression output {name!r}")
    st = _lib.RegOutputs()
    for name, t in outs.items():
        setattr(st, name, t.data_ptr())
    check(
        lib().fb200_bma_reduce_reg(
            _ptr(outputs, torch.float32),
            _ptr(log_temp, torch.float32, True),
            _ptr(targets, torch.float32, True),
            S,
            B,
            d,
            C.byref(st),
            _stream(),
        ),
        "fb200_bma_reduce_reg",
    )
    return outs


# ------------------------------------------------------------------------------------------- scoring
def aps_scores(probs, targets=None, want_perm=False):
    B, Cn = probs.shape
    scores = torch.empty(B if targets is not None else (B, Cn), dtype=torch.float32, device=probs.device)
    perm = torch.empty((B, Cn), dtype=torch.int32, device=probs.device) if want_perm else None
    check(
        lib().fb200_aps_scores(
            _ptr(probs, torch.float32), _ptr(targets, torch.int32, True), B, Cn, _ptr(scores), _ptr(perm, None, True), _stream()
        ),
        "fb200_aps_scores",
    )
    return (scores, perm) if want_perm else scores


def aps_conformal_sets(probs, val_sorted, threshold, want_scores=False):
    """APS all-class scores + rank-count conformal sets in one kernel: (mask uint8 [B,C], sizes int32 [B][, scores])."""
    B, Cn = probs.shape
    mask = torch.empty((B, Cn), dtype=torch.uint8, device=probs.device)
    sizes = torch.empty(B, dtype=torch.int32, device=probs.device)
    scores = torch.empty((B, Cn), dtype=torch.float32, device=probs.device) if want_scores else None
    check(
        lib().fb200_aps_conformal_sets(_ptr(probs, torch.float32), B, Cn, _ptr(val_sorted, torch.float32), val_sorted.numel(),
                                       float(np.float32(threshold)), _ptr(scores, torch.float32, True), _ptr(mask), _ptr(sizes),
                                       _stream()),
        "fb200_aps_conformal_sets",
    )
    return (mask, sizes, scores) if want_scores else (mask, sizes)


def simple_scores(probs, targets=None):
    B, Cn = probs.shape
    scores = torch.empty(B if targets is not None else (B, Cn), dtype=torch.float32, device=probs.device)
    check(
        lib().fb200_simple_scores(_ptr(probs, torch.float32), _ptr(targets, torch.int32, True), B, Cn, _ptr(scores), _stream()),
        "fb200_simple_scores",
    )
    return scores


def credible_sets(mean, error):
    B, Cn = mean.shape
    labels = torch.empty((B, Cn), dtype=torch.int32, device=mean.device)
    size = torch.empty(B, dtype=torch.int32, device=mean.device)
    check(
        lib().fb200_credible_sets(
            _ptr(mean, torch.float32), B, Cn, float(np.float32(1 - error)), _ptr(labels), _ptr(size), _stream()
        ),
        "fb200_credible_sets",
    )
    return labels, size


def sort_f32_(x):
    """In-place ascending sort of a 1-D float32 CUDA tensor."""
    n = x.numel()
    nbytes = lib().fb200_sort_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    check(lib().fb200_sort_f32(_ptr(x, torch.float32), n, _ptr(ws), nbytes, _stream()), "fb200_sort_f32")
    return x


def conformal_sets(val_scores_sorted, test_scores, threshold):
    B, Cn = test_scores.shape
    mask = torch.empty((B, Cn), dtype=torch.uint8, device=test_scores.device)
    sizes = torch.empty(B, dtype=torch.int32, device=test_scores.device)
    check(
        lib().fb200_conformal_sets(
            _ptr(val_scores_sorted, torch.float32),
            val_scores_sorted.numel(),
            _ptr(test_scores, torch.float32),
            B,
            Cn,
            float(np.float32(threshold)),
            _ptr(mask),
            _ptr(sizes),
            _stream(),
        ),
        "fb200_conformal_sets",
    )
    return mask, sizes


def ece_bins_from_rows(confidence, preds, targets, thresholds, brier_terms=None, want_bin_idx=False):
    """ece_bins from the per-row confidence / prediction / Brier term the reduction kernel emits (no pass over [B,C])."""
    B = confidence.shape[0]
    nb = thresholds.numel()
    dev = confidence.device
    counts = torch.empty(nb, dtype=torch.int32, device=dev)
    correct = torch.empty(nb, dtype=torch.int32, device=dev)
    conf_sums = torch.empty(nb, dtype=torch.float32, device=dev)
    result = torch.empty(4 + 2 * nb, dtype=torch.float32, device=dev)
    bin_idx = torch.empty(B, dtype=torch.int32, device=dev) if want_bin_idx else None
    check(
        lib().fb200_ece_bins_from_rows(
            _ptr(confidence, torch.float32), _ptr(preds, torch.int32), _ptr(brier_terms, torch.float32, True),
            _ptr(targets, torch.int32, True), B, _ptr(thresholds, torch.float32), nb, _ptr(bin_idx, None, True),
            _ptr(counts), _ptr(conf_sums), _ptr(correct), _ptr(result), _stream(),
        ),
        "fb200_ece_bins_from_rows",
    )
    return {"counts": counts, "conf_sums": conf_sums, "correct": correct, "result": result[:4], "bin_idx": bin_idx,
            "confs": result[4 : 4 + nb], "accs": result[4 + nb : 4 + 2 * nb]}


def ece_bins(probs, targets, thresholds, preds=None, want_bin_idx=False):
    if probs.dim() == 1:
        B, Cn = probs.shape[0], 1
    else:
        B, Cn = probs.shape
    nb = thresholds.numel()
    dev = probs.device
    counts = torch.empty(nb, dtype=torch.int32, device=dev)
    correct = torch.empty(nb, dtype=torch.int32, device=dev)
    conf_sums = torch.empty(nb, dtype=torch.float32, device=dev)
    result = torch.empty(4 + 2 * nb, dtype=torch.float32, device=dev)
    bin_idx = torch.empty(B, dtype=torch.int32, device=dev) if want_bin_idx else None
    check(
        lib().fb200_ece_bins(
            _ptr(probs, torch.float32),
            B,
            Cn,
            _ptr(preds, torch.int32, True),
            _ptr(targets, torch.int32, True),
            _ptr(thresholds, torch.float32),
            nb,
            _ptr(bin_idx, None, True),
            _ptr(counts),
            _ptr(conf_sums),
            _ptr(correct),
            _ptr(result),
            _stream(),
        ),
        "fb200_ece_bins",
    )
    return {"counts": counts, "conf_sums": conf_sums, "correct": correct, "result": result[:4], "bin_idx": bin_idx,
            "confs": result[4 : 4 + nb], "accs": result[4 + nb : 4 + 2 * nb]}


def ece_from_bins(counts, conf_sums, correct, n_total):
    """(ECE, MCE) float32[2] from additive bin counters (after an all-reduce over shards)."""
    out = torch.empty(2, dtype=torch.float32, device=counts.device)
    check(
        lib().fb200_ece_from_bins(_ptr(counts, torch.int32), _ptr(conf_sums, torch.float32), _ptr(correct, torch.int32),
                                  counts.numel(), int(n_total), _ptr(out), _stream()),
        "fb200_ece_from_bins",
    )
    return out


def ece_pack_bins(ece, out=None, accumulate=False):
    """float64 [3 n_bins] = [counts | correct | conf_sums] of an ece_bins result: ONE all-reduce for a sharded evaluation;
    ``accumulate`` adds to ``out`` (bins of a loader walked batch by batch)."""
    nb = ece["counts"].numel()
    if accumulate and out is None:
        raise ValueError("accumulate needs the running `out`")
    packed = out if out is not None else torch.empty(3 * nb, dtype=torch.float64, device=ece["counts"].device)
    check(lib().fb200_ece_pack_bins(_ptr(ece["counts"], torch.int32), _ptr(ece["conf_sums"], torch.float32),
                                    _ptr(ece["correct"], torch.int32), nb, _ptr(packed, torch.float64), int(bool(accumulate)),
                                    _stream()),
          "fb200_ece_pack_bins")
    return packed


def ece_from_packed_bins(packed, n_total, want_counts=False):
    """(ECE, MCE) float32[2] (and the int32 bin counts) from the all-reduced packed counters."""
    nb = packed.numel() // 3
    out = torch.empty(2, dtype=torch.float32, device=packed.device)
    counts = torch.empty(nb, dtype=torch.int32, device=packed.device) if want_counts else None
    check(lib().fb200_ece_from_packed_bins(_ptr(packed, torch.float64), nb, int(n_total), _ptr(counts, torch.int32, True),
                                           _ptr(out), _stream()), "fb200_ece_from_packed_bins")
    return (out, counts) if want_counts else out


def quantile_sorted(x_sorted, q):
    out = torch.empty(q.numel(), dtype=torch.float32, device=x_sorted.device)
    check(
        lib().fb200_quantile_sorted(_ptr(x_sorted, torch.float32), x_sorted.numel(), _ptr(q, torch.float32), q.numel(), _ptr(out), _stream()),
        "fb200_quantile_sorted",
    )
    return out


def quantile_axis0(x, q):
    """jnp.quantile(x, q, axis=0): x [T, ...] float32 -> [nq, ...]."""
    T = x.shape[0]
    M = x.numel() // T
    out = torch.empty((q.numel(),) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
    check(
        lib().fb200_quantile_axis0(_ptr(x, torch.float32), T, M, _ptr(q, torch.float32), q.numel(), _ptr(out), _stream()),
        "fb200_quantile_axis0",
    )
    return out


def reg_sample_targets(outputs, key, n_target_samples, log_temp=None, want_idx=False):
    """outputs [n_stored, B, 2d] -> y [T, B, d] (and the drawn posterior-sample indices int32[T])."""
    n, B, two_d = outputs.shape
    d = two_d // 2
    y = torch.empty((n_target_samples, B, d), dtype=torch.float32, device=outputs.device)
    idx = torch.empty(n_target_samples, dtype=torch.int32, device=outputs.device) if want_idx else None
    check(
        lib().fb200_reg_sample_targets(_ptr(outputs, torch.float32), _ptr(log_temp, torch.float32, True), _ptr(key, KEY_DTYPE),
                                       int(n_target_samples), n, B, d, _ptr(y), _ptr(idx, None, True), _stream()),
        "fb200_reg_sample_targets",
    )
    return (y, idx) if want_idx else y


def calib_cls_loss_terms(logits, targets, log_temp=None):
    """logits [S,B,C] (uncalibrated), targets int32 [B], log_temp float32 [1] -> float64 [2, S]:
    row 0 = sum_b log softmax(exp(-log_temp) z[s,b])[y_b], row 1 = its derivative with respect to exp(-log_temp)."""
    S, B, Cn = logits.shape
    nbytes = int(lib().fb200_calib_cls_workspace_bytes(S, B))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=logits.device)
    out = torch.empty((2, S), dtype=torch.float64, device=logits.device)
    check(
        lib().fb200_calib_cls_loss_terms(_ptr(logits, (torch.float32, torch.bfloat16)), _DT[logits.dtype],
                                         _ptr(targets, torch.int32), _ptr(log_temp, torch.float32, True), S, B, Cn,
                                         _ptr(ws), nbytes, _ptr(out), _stream()),
        "fb200_calib_cls_loss_terms",
    )
    return out


def calib_reg_loss_terms(outputs, targets, log_temp=None):
    """outputs [S,B,2d] (uncalibrated [mu | log var]), targets float32 [B,d] -> float64 [2, S]: Gaussian log-likelihood sums
    with log var + log_temp and their derivative in log_temp."""
    S, B, two_d = outputs.shape
    nbytes = int(lib().fb200_calib_reg_workspace_bytes(S, B))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=outputs.device)
    out = torch.empty((2, S), dtype=torch.float64, device=outputs.device)
    check(
        lib().fb200_calib_reg_loss_terms(_ptr(outputs, torch.float32), _ptr(targets, torch.float32),
                                         _ptr(log_temp, torch.float32, True), S, B, two_d // 2, _ptr(ws), nbytes,
                                         _ptr(out), _stream()),
        "fb200_calib_reg_loss_terms",
    )
    return out


def output_calib_focal_loss_terms(logits, targets, log_temp=None, gamma=2.0):
    """logits [B,C] (uncalibrated), int32 targets [B] -> float64 [2]: sum_b -(1-p_b)^gamma log p_b with
    p_b = softmax(exp(-log_temp) z_b)[y_b], and its derivative with respect to exp(-log_temp)."""
    B, Cn = logits.shape
    nbytes = int(lib().fb200_calib_cls_workspace_bytes(1, B))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=logits.device)
    out = torch.empty(2, dtype=torch.float64, device=logits.device)
    check(
        lib().fb200_output_calib_focal_loss_terms(_ptr(logits, (torch.float32, torch.bfloat16)), _DT[logits.dtype],
                                                  _ptr(targets, torch.int32), _ptr(log_temp, torch.float32, True), B, Cn,
                                                  float(gamma), _ptr(ws), nbytes, _ptr(out), _stream()),
        "fb200_output_calib_focal_loss_terms",
    )
    return out


def output_calib_scaled_mse_terms(outputs, targets, log_temp=None):
    """outputs [B,2d] (uncalibrated [mu | log var]), targets float32 [B,d] -> float64 [2]: the variance-scaled squared
    error summed over b with log var + log_temp, and its derivative in log_temp."""
    B, two_d = outputs.shape
    nbytes = int(lib().fb200_calib_reg_workspace_bytes(1, B))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=outputs.device)
    out = torch.empty(2, dtype=torch.float64, device=outputs.device)
    check(
        lib().fb200_output_calib_scaled_mse_terms(_ptr(outputs, torch.float32), _ptr(targets, torch.float32),
                                                  _ptr(log_temp, torch.float32, True), B, two_d // 2, _ptr(ws), nbytes,
                                                  _ptr(out), _stream()),
        "fb200_output_calib_scaled_mse_terms",
    )
    return out


def focal_loss_grad(logits, targets, gamma=2.0, want_grad=True):
    """logits float32 [n,C], targets int32 [n] -> (loss float64 [1] = mean focal loss, grad float32 [n,C] = d loss / d logits
    or None, loss_rows float32 [n]) - the loss head of CalibClassifier.calibrate (fb200_focal_loss_grad)."""
    n, Cn = logits.shape
    grad = torch.empty_like(logits) if want_grad else None
    rows = torch.empty(n, dtype=torch.float32, device=logits.device)
    loss = torch.empty(1, dtype=torch.float64, device=logits.device)
    check(
        lib().fb200_focal_loss_grad(_ptr(logits, torch.float32), _ptr(targets, torch.int32), n, Cn, float(gamma),
                                    _ptr(grad, torch.float32, True), _ptr(rows), _ptr(loss), _stream()),
        "fb200_focal_loss_grad",
    )
    return loss, grad, rows


def scaled_mse_loss_grad(outputs, targets, want_grad=True):
    """outputs float32 [n,2d] = [mu | log var], targets float32 [n,d] -> (loss float64 [1] = mean over n of the summed
    variance-scaled squared errors, grad float32 [n,2d] or None, loss_terms float32 [n,d]) (fb200_scaled_mse_loss_grad)."""
    n, two_d = outputs.shape
    d = two_d // 2
    grad = torch.empty_like(outputs) if want_grad else None
    terms = torch.empty((n, d), dtype=torch.float32, device=outputs.device)
    loss = torch.empty(1, dtype=torch.float64, device=outputs.device)
    check(
        lib().fb200_scaled_mse_loss_grad(_ptr(outputs, torch.float32), _ptr(targets, torch.float32), n, d,
                                         _ptr(grad, torch.float32, True), _ptr(terms), _ptr(loss), _stream()),
        "fb200_scaled_mse_loss_grad",
    )
    return loss, grad, terms


def dot_f32(a, b):
    """float64 [1] = sum a * b over two float32 tensors of equal size, fixed summation order (fb200_dot_f32)."""
    out = torch.empty(1, dtype=torch.float64, device=a.device)
    check(lib().fb200_dot_f32(_ptr(a, torch.float32), _ptr(b, torch.float32), a.numel(), _ptr(out), _stream()), "fb200_dot_f32")
    return out


def adam_step_(params, grads, mu, nu, count, lr=1e-2, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adam step number ``count`` (1-based) over flat float32 vectors, in place (fb200_adam_step_f32)."""
    import numpy as np

    f = np.float32
    c1, c2 = f(1) - f(b1) ** f(count), f(1) - f(b2) ** f(count)
    check(
        lib().fb200_adam_step_f32(_ptr(params, torch.float32), _ptr(grads, torch.float32), _ptr(mu, torch.float32),
                                  _ptr(nu, torch.float32), params.numel(), float(lr), float(b1), float(b2), float(eps),
                                  float(c1), float(c2), _stream()),
        "fb200_adam_step_f32",
    )
    return params


# ------------------------------------------------------------------------------------------- multivalid scorers
def multivalid_stats(values, scores, groups, buckets, top_label=False, mode=0, taus=(0, 1, 2)):
    """values float32 [n, Dv], scores float32 [n, D], groups uint8 [n, G] or None, buckets float32 [n_buckets] ->
    float64 [3, G, n_buckets, D, 3]: per (bucket type, group, bucket, dimension) the member count, the sum of the
    statistic (mode 0: the score; mode 1: [score <= value]) and the sum of the values over the members
    (fb200_multivalid_stats)."""
    n, Dv = values.shape
    D = scores.shape[1]
    G = 1 if groups is None else groups.shape[1]
    nb = buckets.numel()
    nbytes = int(lib().fb200_multivalid_workspace_bytes(n, G, nb, D))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=values.device)
    out = torch.empty((3, G, nb, D, 3), dtype=torch.float64, device=values.device)
    mask = sum(1 << int(t) for t in taus)
    check(
        lib().fb200_multivalid_stats(_ptr(values, torch.float32), _ptr(scores, torch.float32), _ptr(groups, torch.uint8, True),
                                     _ptr(buckets, torch.float32), n, D, Dv, G, nb, int(bool(top_label)), int(mode), mask,
                                     _ptr(ws), nbytes, _ptr(out), _stream()),
        "fb200_multivalid_stats",
    )
    return out


def multivalid_patch_(values, groups, g, c, tau, v, n_buckets, eta, patch, top_label=False, multiplicative=False):
    """In place: the (tau, g, v, c) members of values [n, Dv] move by the patch and are clipped to [0, 1]; returns the
    member count as an int32 [1] tensor (fb200_multivalid_patch)."""
    n, Dv = values.shape
    G = 1 if groups is None else groups.shape[1]
    cnt = torch.empty(1, dtype=torch.int32, device=values.device)
    check(
        lib().fb200_multivalid_patch(_ptr(values, torch.float32), _ptr(groups, torch.uint8, True), n, Dv, G, int(g), int(c),
                                     int(tau), float(v), int(n_buckets), int(bool(top_label)), float(eta), float(patch),
                                     int(bool(multiplicative)), _ptr(cnt), _stream()),
        "fb200_multivalid_patch",
    )
    return cnt


def multivalid_loss(values, scores, mode=0, coverage=0.0):
    """float64 [1]: sum over all elements of (values - scores)^2 (mode 0) or of the pinball terms at ``coverage`` (mode 1);
    values float32 [n, Dv], scores float32 [n, D], Dv = 1 (scores[:, 0] is used) or D (fb200_multivalid_loss)."""
    n, Dv = values.shape
    D = scores.shape[1]
    terms = torch.empty(n * Dv, dtype=torch.float32, device=values.device)
    out = torch.empty(1, dtype=torch.float64, device=values.device)
    check(
        lib().fb200_multivalid_loss(_ptr(values, torch.float32), _ptr(scores, torch.float32), n, D, Dv, int(mode),
                                    float(coverage), _ptr(terms), _ptr(out), _stream()),
        "fb200_multivalid_loss",
    )
    return out


def batchmvp_delta_counts(values, scores, groups, g, tau, v, n_buckets, deltas):
    """int64 [m + 1]: per delta the members of (tau, g, v) with score <= value + delta; last entry the member count
    (fb200_batchmvp_delta_counts).  values / scores float32 [n]."""
    n = values.numel()
    G = 1 if groups is None else groups.shape[1]
    m = deltas.numel()
    counts = torch.empty(m + 1, dtype=torch.int64, device=values.device)
    check(
        lib().fb200_batchmvp_delta_counts(_ptr(values, torch.float32), _ptr(scores, torch.float32),
                                          _ptr(groups, torch.uint8, True), n, G, int(g), int(tau), float(v), int(n_buckets),
                                          _ptr(deltas, torch.float32), m, _ptr(counts), _stream()),
        "fb200_batchmvp_delta_counts",
    )
    return counts


def maxcov_counts(probs, targets, taus, side, threshold):
    """probs float32 [n], targets int32 [n], taus float32 [n_taus] sorted (ascending for side 0, descending for side 1)
    -> (counts, labels) int32 [n_taus]: per candidate tau the number of selected points and the sum of their counted
    labels (fb200_maxcov_counts)."""
    n, nt = probs.numel(), taus.numel()
    nbytes = int(lib().fb200_maxcov_workspace_bytes(nt))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=probs.device)
    counts = torch.empty(nt, dtype=torch.int32, device=probs.device)
    labels = torch.empty(nt, dtype=torch.int32, device=probs.device)
    check(
        lib().fb200_maxcov_counts(_ptr(probs, torch.float32), _ptr(targets, torch.int32), n, _ptr(taus, torch.float32), nt,
                                  int(side), float(np.float32(threshold)), _ptr(ws), nbytes, _ptr(counts), _ptr(labels),
                                  _stream()),
        "fb200_maxcov_counts",
    )
    return counts, labels


def maxcov_apply_patches(probs, tau_pos, tau_neg):
    out = torch.empty_like(probs)
    check(
        lib().fb200_maxcov_apply_patches(_ptr(probs, torch.float32), probs.numel(), float(np.float32(tau_pos)),
                                         float(np.float32(tau_neg)), _ptr(out), _stream()),
        "fb200_maxcov_apply_patches",
    )
    return out


def cls_output_sample_targets(logits, key, n_target_samples, log_temp=None):
    """logits [B, C] -> y int32 [T, B]: ClassificationProbOutputLayer.sample's draws (T uniforms per input key)."""
    B, Cn = logits.shape
    y = torch.empty((n_target_samples, B), dtype=torch.int32, device=logits.device)
    check(
        lib().fb200_cls_output_sample_targets(_ptr(logits, torch.float32), _ptr(log_temp, torch.float32, True),
                                              _ptr(key, KEY_DTYPE), int(n_target_samples), B, Cn, _ptr(y), _stream()),
        "fb200_cls_output_sample_targets",
    )
    return y


def reg_output_sample_targets(outputs, key, n_target_samples, log_temp=None):
    """outputs [B, 2d] -> y float32 [T, B, d] = mu + exp(0.5 (log var + log_temp)) * normal(key, (T, B, d))."""
    B, two_d = outputs.shape
    d = two_d // 2
    y = torch.empty((n_target_samples, B, d), dtype=torch.float32, device=outputs.device)
    check(
        lib().fb200_reg_output_sample_targets(_ptr(outputs, torch.float32), _ptr(log_temp, torch.float32, True),
                                              _ptr(key, KEY_DTYPE), int(n_target_samples), B, d, _ptr(y), _stream()),
        "fb200_reg_output_sample_targets",
    )
    return y


def cls_sample_targets(logits, key, n_target_samples, log_temp=None, want_idx=False):
    """logits [n_stored, B, C] -> y int32 [T, B] (and the drawn posterior-sample indices int32 [T])."""
    n, B, Cn = logits.shape
    y = torch.empty((n_target_samples, B), dtype=torch.int32, device=logits.device)
    idx = torch.empty(n_target_samples, dtype=torch.int32, device=logits.device) if want_idx else None
    check(
        lib().fb200_cls_sample_targets(_ptr(logits, torch.float32), _ptr(log_temp, torch.float32, True), _ptr(key, KEY_DTYPE),
                                       int(n_target_samples), n, B, Cn, _ptr(y), _ptr(idx, None, True), _stream()),
        "fb200_cls_sample_targets",
    )
    return (y, idx) if want_idx else y


def reg_mc_entropies(outputs, sample_idx, target_keys, n_target_samples, names=("entropy",), log_temp=None):
    """outputs [n_stored, B, 2d], sample_idx int32 [S], target_keys int32 [S, 2] -> dict name -> float32 [B] for names
    among entropy / aleatoric_entropy / epistemic_entropy."""
    n, B, two_d = outputs.shape
    S = sample_idx.numel()
    res = {k: torch.empty(B, dtype=torch.float32, device=outputs.device) for k in names}
    for k in names:
        if k not in ("entropy", "aleatoric_entropy", "epistemic_entropy"):
            raise ValueError(k)
    check(
        lib().fb200_reg_mc_entropies(
            _ptr(outputs, torch.float32), _ptr(log_temp, torch.float32, True), _ptr(sample_idx, torch.int32),
            _ptr(target_keys, KEY_DTYPE), n, S, int(n_target_samples), B, two_d // 2,
            _ptr(res.get("entropy"), None, True), _ptr(res.get("aleatoric_entropy"), None, True),
            _ptr(res.get("epistemic_entropy"), None, True), _stream(),
        ),
        "fb200_reg_mc_entropies",
    )
    return res


# ------------------------------------------------------------------------------------------------ diagnostics
def _row_matrix(t):
    """[S, N] float32 CUDA matrix whose rows may be strided (a slice of the sample bank): (ptr, S, N, row_stride)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2:
        raise ValueError("expected a 2-D float32 CUDA tensor [n_samples, n_params]")
    if t.shape[1] > 1 and t.stride(1) != 1:
        t = t.contiguous()
    S, N = t.shape
    stride = t.stride(0) if S > 1 else N
    if stride < N:
        t = t.contiguous()
        stride = N
    return t, C.c_void_p(t.data_ptr()), int(S), int(N), int(stride)


def ksd_imq(samples, grads, c=1.0, beta=-0.5):
    """float32[1] device tensor: kernel Stein discrepancy with the IMQ kernel over S samples x N parameters."""
    xs, xp, S, N, sx = _row_matrix(samples)
    gs, gp, S2, N2, sg = _row_matrix(grads)
    if (S, N) != (S2, N2):
        raise ValueError("samples and grads must have the same shape")
    if sx != sg:
        xs, gs = xs.contiguous(), gs.contiguous()
        xp, gp, sx = C.c_void_p(xs.data_ptr()), C.c_void_p(gs.data_ptr()), N
    nbytes = lib().fb200_ksd_imq_workspace_bytes(S, N)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=xs.device)
    out = torch.empty(1, dtype=torch.float32, device=xs.device)
    check(
        lib().fb200_ksd_imq(xp, gp, S, N, sx, float(c), float(beta), _ptr(ws), nbytes, _ptr(out), _stream()),
        "fb200_ksd_imq",
    )
    return out


def ess(samples, filter_threshold=0.0):
    """float32[N]: per-parameter effective sample size of the chain samples[S, N]."""
    xs, xp, S, N, sx = _row_matrix(samples)
    out = torch.empty(N, dtype=torch.float32, device=xs.device)
    use = filter_threshold is not None
    check(
        lib().fb200_ess(xp, S, N, sx, int(use), float(filter_threshold if use else 0.0), _ptr(out), _stream()),
        "fb200_ess",
    )
    return out


# ------------------------------------------------------------------------------------------------ conformal Bayes
def conformal_bayes_sets(ell_train, lsm_test, error, want_counts=False, want_ess=False):
    """ell_train [S, n_train], lsm_test [S, n_test, C] (float32) -> dict(region uint8 [n_test, C], sizes int32 [n_test],
    optionally rank_counts int32 [n_test, C] and ess float32 [n_test, C])."""
    S, n_train = ell_train.shape
    S2, n_test, Cn = lsm_test.shape
    if S != S2:
        raise ValueError("training and test log-probabilities must come from the same posterior draws")
    dev = ell_train.device
    thr = float(np.float32(error * (n_train + 1)))
    nbytes = int(lib().fb200_conformal_bayes_workspace_bytes(S, n_test, Cn))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    region = torch.empty((n_test, Cn), dtype=torch.uint8, device=dev)
    sizes = torch.empty(n_test, dtype=torch.int32, device=dev)
    counts = torch.empty((n_test, Cn), dtype=torch.int32, device=dev) if want_counts else None
    ess_t = torch.empty((n_test, Cn), dtype=torch.float32, device=dev) if want_ess else None
    check(
        lib().fb200_conformal_bayes_sets(
            _ptr(ell_train, torch.float32), _ptr(lsm_test, torch.float32), S, n_train, n_test, Cn, thr, _ptr(ws), nbytes,
            _ptr(region), _ptr(sizes), _ptr(counts, torch.int32, True), _ptr(ess_t, torch.float32, True), _stream(),
        ),
        "fb200_conformal_bayes_sets",
    )
    out = {"region": region, "sizes": sizes}
    if want_counts:
        out["rank_counts"] = counts
    if want_ess:
        out["ess"] = ess_t
    return out


def softmax_rows_(z, epilogue=EPI_LOG_SOFTMAX, want_lse=False, log_temp=None):
    """In place over the last axis of a float32 / bf16 CUDA tensor: softmax or log_softmax of z * exp(-log_temp) (or
    EPI_LOGITS: the calibrated logits themselves, plus the row logsumexp) - the row pass fb200_last_layer_logits_ex uses
    off the tcgen05 path."""
    Cn = z.shape[-1]
    rows = z.numel() // Cn
    lse = torch.empty(z.shape[:-1], dtype=torch.float32, device=z.device) if want_lse else None
    check(
        lib().fb200_softmax_rows(_ptr(z, (torch.float32, torch.bfloat16)), _DT[z.dtype], rows, Cn, int(epilogue),
                                 _ptr(log_temp, torch.float32, True), _ptr(lse, torch.float32, True), _stream()),
        "fb200_softmax_rows",
    )
    return (z, lse) if want_lse else z
