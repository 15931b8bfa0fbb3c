"""Out-of-bounds writes: every output (and workspace) of the newer kernels is carved out of a larger canary-filled
buffer; after the call the guard zones on both sides must be untouched.  (compute-sanitizer is not available on the
GPU pool, so this is the bounds check.)  Shapes are chosen ragged on purpose: sizes not divisible by the tile sizes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 64  # elements on each side


class Guarded:
    def __init__(self, shape, dtype, device):
        n = int(np.prod(shape))
        self.canary = {torch.float32: 1234.5, torch.float64: 1234.5, torch.int32: 0x5A5A5A5, torch.uint8: 0xA5}[dtype]
        self.buf = torch.full((n + 2 * GUARD,), self.canary, dtype=dtype, device=device)
        self.view = self.buf[GUARD : GUARD + n].view(shape)
        self.n = n

    def intact(self):
        lo, hi = self.buf[:GUARD], self.buf[GUARD + self.n :]
        return bool((lo == self.canary).all()) and bool((hi == self.canary).all())


def _call(lib_fn, name, *args):
    from fortuna_b200._lib import check

    check(lib_fn(*args), name)
    torch.cuda.synchronize()


def test_guards_scoring_and_sampling(cuda):
    from fortuna_b200 import ops
    from fortuna_b200._lib import lib
    from fortuna_b200.ops import _ptr, _stream

    r = np.random.default_rng(0)
    B, Cn, n_val = 37, 1000, 123
    probs = torch.softmax(torch.from_numpy((3 * r.standard_normal((B, Cn))).astype(np.float32)).to(cuda), -1).contiguous()
    val = ops.sort_f32_(torch.rand(n_val, device=cuda))
    mask, sizes, scores = Guarded((B, Cn), torch.uint8, cuda), Guarded((B,), torch.int32, cuda), Guarded((B, Cn), torch.float32, cuda)
    _call(lib().fb200_aps_conformal_sets, "fused", _ptr(probs), B, Cn, _ptr(val), n_val, float(np.float32(0.9 * (n_val + 1))),
          _ptr(scores.view), _ptr(mask.view), _ptr(sizes.view), _stream())
    assert mask.intact() and sizes.intact() and scores.intact()
    assert int(mask.view.max()) <= 1 and int(sizes.view.min()) >= 0
    # class samples: [T, B] int32
    z = torch.from_numpy((2 * r.standard_normal((3, B, 37))).astype(np.float32)).to(cuda)
    y, idx = Guarded((5, B), torch.int32, cuda), Guarded((5,), torch.int32, cuda)
    _call(lib().fb200_cls_sample_targets, "cls_sample", _ptr(z), None, _ptr(ops.PRNGKey(1, cuda)), 5, 3, B, 37, _ptr(y.view),
          _ptr(idx.view), _stream())
    assert y.intact() and idx.intact() and int(y.view.max()) < 37 and int(y.view.min()) >= 0
    # conformal Bayes
    S, n_train, n_test, Cc = 7, 65, 9, 3
    ell = torch.log_softmax(torch.randn((S, n_train, Cc), device=cuda), -1)[..., 0].contiguous()
    lsm = torch.log_softmax(torch.randn((S, n_test, Cc), device=cuda), -1).contiguous()
    nbytes = int(lib().fb200_conformal_bayes_workspace_bytes(S, n_test, Cc))
    ws = Guarded((nbytes,), torch.uint8, cuda)
    reg, sz, cnt, ess = (Guarded((n_test, Cc), torch.uint8, cuda), Guarded((n_test,), torch.int32, cuda),
                         Guarded((n_test, Cc), torch.int32, cuda), Guarded((n_test, Cc), torch.float32, cuda))
    _call(lib().fb200_conformal_bayes_sets, "cb", _ptr(ell), _ptr(lsm), S, n_train, n_test, Cc, 3.3, _ptr(ws.view), nbytes,
          _ptr(reg.view), _ptr(sz.view), _ptr(cnt.view), _ptr(ess.view), _stream())
    assert reg.intact() and sz.intact() and cnt.intact() and ess.intact()
    assert ws.intact()


def test_guards_reductions_and_diagnostics(cuda):
    from fortuna_b200 import _lib as L
    from fortuna_b200 import ops
    from fortuna_b200._lib import lib
    from fortuna_b200.ops import _ptr, _stream

    r = np.random.default_rng(1)
    # narrow and wide reductions: [B,C] / [B] / [S,B] outputs
    for S, B, Cn in ((5, 33, 10), (3, 7, 1300)):
        z = torch.from_numpy((2 * r.standard_normal((S, B, Cn))).astype(np.float32)).to(cuda)
        t = torch.from_numpy(r.integers(0, Cn, B).astype(np.int32)).to(cuda)
        g = {"mean": Guarded((B, Cn), torch.float32, cuda), "epistemic_variance": Guarded((B, Cn), torch.float32, cuda),
             "entropy": Guarded((B,), torch.float32, cuda), "log_prob": Guarded((B,), torch.float32, cuda),
             "mode": Guarded((B,), torch.int32, cuda), "ensemble_log_prob": Guarded((S, B), torch.float32, cuda)}
        # [B,C] outputs must be 16-byte aligned for the vector epilogue: GUARD elements of 4 bytes keep that
        out = ops.bma_reduce_cls(z, tuple(g), targets=t, out={k: v.view for k, v in g.items()})
        torch.cuda.synchronize()
        assert all(v.intact() for v in g.values()), (S, B, Cn)
        assert bool(torch.isfinite(out["mean"]).all())
    # regression Monte-Carlo entropies and calibration terms
    zr = torch.randn((4, 19, 6), device=cuda)
    ks = ops.split(ops.PRNGKey(2, cuda), 6)
    idx = ops.sample_indices(ks[0].contiguous(), 5, 4)
    e = [Guarded((19,), torch.float32, cuda) for _ in range(3)]
    _call(lib().fb200_reg_mc_entropies, "reg_ent", _ptr(zr), None, _ptr(idx), _ptr(ks[1:].contiguous()), 4, 5, 3, 19, 3,
          _ptr(e[0].view), _ptr(e[1].view), _ptr(e[2].view), _stream())
    assert all(x.intact() for x in e)
    zc = torch.randn((6, 45, 11), device=cuda)
    tc = torch.randint(0, 11, (45,), device=cuda, dtype=torch.int32)
    nb = int(lib().fb200_calib_cls_workspace_bytes(6, 45))
    ws, out = Guarded((nb // 8,), torch.float64, cuda), Guarded((12,), torch.float64, cuda)
    _call(lib().fb200_calib_cls_loss_terms, "calib", _ptr(zc), L.F32, _ptr(tc), None, 6, 45, 11, _ptr(ws.view), nb, _ptr(out.view), _stream())
    assert ws.intact() and out.intact()
    # diagnostics: ESS output and the KSD workspace
    x, gq = torch.randn((33, 70), device=cuda), torch.randn((33, 70), device=cuda)
    ess = Guarded((70,), torch.float32, cuda)
    _call(lib().fb200_ess, "ess", _ptr(x), 33, 70, 70, 1, 0.0, _ptr(ess.view), _stream())
    assert ess.intact()
    nb = int(lib().fb200_ksd_imq_workspace_bytes(33, 70))
    ws, out = Guarded((nb // 4,), torch.float32, cuda), Guarded((1,), torch.float32, cuda)
    _call(lib().fb200_ksd_imq, "ksd", _ptr(x), _ptr(gq), 33, 70, 70, 1.0, -0.5, _ptr(ws.view), nb, _ptr(out.view), _stream())
    assert ws.intact() and out.intact() and bool(torch.isfinite(out.view).all())
