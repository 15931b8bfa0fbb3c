"""K2 tensor-core path (tcgen05 / TMEM / TMA) vs the oracle's bf16-input, wide-accumulate GEMM.

north_star tolerance for the bf16 GEMM is rtol 2e-2; against an oracle fed the SAME bf16-rounded inputs
the only difference is float32 accumulation order, so the test holds it to 2e-3 (abs floor scaled by sqrt(D)).
"""
import numpy as np
import pytest
import torch

from oracle import predictive as P

pytestmark = pytest.mark.gpu


def _run(cuda, S, B, D, C, out_dtype=torch.float32, with_bias=True, with_temp=True, seed=0):
    from fortuna_b200 import ops

    r = np.random.default_rng(seed)
    x = r.standard_normal((B, D)).astype(np.float32)
    w = (r.standard_normal((S, D, C)) / np.sqrt(D)).astype(np.float32)
    bias = r.standard_normal((S, C)).astype(np.float32) if with_bias else None
    lt = np.array([0.3], dtype=np.float32) if with_temp else None
    want = P.last_layer_logits(x, w, bias, lt, bf16_inputs=True)
    xt = torch.from_numpy(x).to(cuda).to(torch.bfloat16)
    wt = ops.transpose_w_bf16(torch.from_numpy(w).to(cuda))  # [S, C, D] bf16, K-major
    np.testing.assert_array_equal(
        wt.float().cpu().numpy(), np.transpose(P.round_to_bf16(w), (0, 2, 1)), err_msg="transpose_w_bf16"
    )
    out = ops.last_layer_logits(
        xt, wt,
        torch.from_numpy(bias).to(cuda) if with_bias else None,
        torch.from_numpy(lt).to(cuda) if with_temp else None,
        out_dtype=out_dtype, w_layout=ops.W_SCD, path=2,
    )
    torch.cuda.synchronize()
    got = out.float().cpu().numpy()
    tol = 2e-3 if out_dtype == torch.float32 else 1e-2
    np.testing.assert_allclose(got, want, rtol=tol, atol=tol * np.sqrt(D) / 8)
    return got, want


@pytest.mark.parametrize(
    "S,B,D,C",
    [(1, 128, 64, 256), (2, 300, 128, 1000), (3, 129, 512, 10), (2, 1000, 2048, 1000), (1, 77, 72, 40),
     (5, 4096, 512, 1000)],
)
def test_tc_matches_oracle_f32_out(cuda, S, B, D, C):
    _run(cuda, S, B, D, C)


def test_tc_bf16_out_no_bias_no_temp(cuda):
    _run(cuda, 2, 256, 256, 512, out_dtype=torch.bfloat16, with_bias=False, with_temp=False)


def test_tc_equals_cuda_core_path(cuda):
    from fortuna_b200 import ops

    r = np.random.default_rng(3)
    S, B, D, C = 2, 200, 192, 300
    xt = torch.from_numpy(r.standard_normal((B, D)).astype(np.float32)).to(cuda).to(torch.bfloat16)
    wt = torch.from_numpy((r.standard_normal((S, C, D)) / 14).astype(np.float32)).to(cuda).to(torch.bfloat16)
    a = ops.last_layer_logits(xt, wt, w_layout=ops.W_SCD, path=1)
    b = ops.last_layer_logits(xt, wt, w_layout=ops.W_SCD, path=2)
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-4, atol=1e-4)


def test_tc_rejects_unsupported(cuda):
    from fortuna_b200 import _lib, ops

    x = torch.zeros((8, 64), dtype=torch.float32, device=cuda)
    w = torch.zeros((1, 64, 16), dtype=torch.float32, device=cuda)
    with pytest.raises(_lib.FortunaB200Error):
        ops.last_layer_logits(x, w, path=2)  # float32 / [S,D,C] layout is not a tensor-core shape


@pytest.mark.parametrize("path", [1, 2])
def test_sample_index_gather_is_fused(cuda, path):
    """out[s] uses w[idx[s]] / bias[idx[s]] (posterior.sample() draws with replacement) on both paths."""
    from fortuna_b200 import ops

    r = np.random.default_rng(9)
    n, S, B, D, C = 5, 8, 150, 128, 72
    x = r.standard_normal((B, D)).astype(np.float32)
    w = (r.standard_normal((n, D, C)) / np.sqrt(D)).astype(np.float32)
    bias = r.standard_normal((n, C)).astype(np.float32)
    idx = np.array([4, 0, 0, 3, 1, 4, 2, 2], dtype=np.int32)
    want = P.last_layer_logits(x, w[idx], bias[idx], None, bf16_inputs=True)
    xt = torch.from_numpy(x).to(cuda).to(torch.bfloat16)
    wt = ops.transpose_w_bf16(torch.from_numpy(w).to(cuda))
    out = ops.last_layer_logits(xt, wt, torch.from_numpy(bias).to(cuda), None, w_layout=ops.W_SCD, path=path,
                                sample_idx=torch.from_numpy(idx).to(cuda))
    assert out.shape == (S, B, C)
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("S,B,D,C,out_dtype", [(64, 1000, 768, 2, torch.float32), (7, 300, 256, 10, torch.float32),
                                               (33, 257, 512, 3, torch.bfloat16), (40, 128, 320, 63, torch.float32)])
def test_narrow_last_layer_folds_samples_into_columns(cuda, S, B, D, C, out_dtype):
    """bf16 operands with C < 64 (BERT classifier 768 x 2, ResNet 512 x 10): path 0 runs the tcgen05 kernel on the
    flattened column axis n = s * C + c and scatters back to [S, B, C]; equals the oracle and the CUDA-core path."""
    from fortuna_b200 import ops

    r = np.random.default_rng(S * 100 + C)
    x = r.standard_normal((B, D)).astype(np.float32)
    w = (r.standard_normal((S, D, C)) / np.sqrt(D)).astype(np.float32)
    bias = r.standard_normal((S, C)).astype(np.float32)
    lt = np.array([0.1], dtype=np.float32)
    want = P.last_layer_logits(x, w, bias, lt, bf16_inputs=True)
    xt = torch.from_numpy(x).to(cuda).to(torch.bfloat16)
    wt = ops.transpose_w_bf16(torch.from_numpy(w).to(cuda))
    args = (xt, wt, torch.from_numpy(bias).to(cuda), torch.from_numpy(lt).to(cuda))
    auto = ops.last_layer_logits(*args, out_dtype=out_dtype, w_layout=ops.W_SCD, path=0)
    simt = ops.last_layer_logits(*args, out_dtype=out_dtype, w_layout=ops.W_SCD, path=1)
    tol = 2e-3 if out_dtype == torch.float32 else 1e-2
    np.testing.assert_allclose(auto.float().cpu().numpy(), want, rtol=tol, atol=tol * np.sqrt(D) / 8)
    np.testing.assert_allclose(simt.float().cpu().numpy(), want, rtol=tol, atol=tol * np.sqrt(D) / 8)


@pytest.mark.parametrize(
    "S,B,D,C,out_dtype",
    [(1, 256, 64, 256, torch.float32), (2, 300, 128, 1000, torch.float32), (3, 129, 512, 10, torch.float32),
     (2, 1000, 2048, 1000, torch.float32), (1, 77, 72, 40, torch.float32), (5, 4096, 512, 1000, torch.bfloat16),
     (3, 2500, 256, 333, torch.float32)],
)
def test_cta_pair_kernel_matches_oracle(cuda, S, B, D, C, out_dtype):
    """path=3: the cta_group::2 kernel (256 x 256 tile per CTA pair) against the oracle and the 1-CTA kernel."""
    from fortuna_b200 import ops

    r = np.random.default_rng(S + B + C)
    x = r.standard_normal((B, D)).astype(np.float32)
    w = (r.standard_normal((S, D, C)) / np.sqrt(D)).astype(np.float32)
    bias = r.standard_normal((S, C)).astype(np.float32)
    lt = np.array([0.3], dtype=np.float32)
    want = P.last_layer_logits(x, w, bias, lt, bf16_inputs=True)
    xt = torch.from_numpy(x).to(cuda).to(torch.bfloat16)
    wt = ops.transpose_w_bf16(torch.from_numpy(w).to(cuda))
    args = (xt, wt, torch.from_numpy(bias).to(cuda), torch.from_numpy(lt).to(cuda))
    pair = ops.last_layer_logits(*args, out_dtype=out_dtype, w_layout=ops.W_SCD, path=3)
    one = ops.last_layer_logits(*args, out_dtype=out_dtype, w_layout=ops.W_SCD, path=2)
    torch.cuda.synchronize()
    tol = 2e-3 if out_dtype == torch.float32 else 1e-2
    np.testing.assert_allclose(pair.float().cpu().numpy(), want, rtol=tol, atol=tol * np.sqrt(D) / 8)
    assert torch.equal(pair, one)  # same K order and fp32 accumulation: bit-identical to the 1-CTA kernel


def _ex_inputs(cuda, S, B, D, C, seed, bf16=True):
    from fortuna_b200 import ops

    r = np.random.default_rng(seed)
    x = r.standard_normal((B, D)).astype(np.float32)
    w = (2.0 * r.standard_normal((S, D, C)) / np.sqrt(D)).astype(np.float32)
    bias = r.standard_normal((S, C)).astype(np.float32)
    lt = np.array([0.2], dtype=np.float32)
    z = P.last_layer_logits(x, w, bias, lt, bf16_inputs=bf16)
    if bf16:
        xt = torch.from_numpy(x).to(cuda).to(torch.bfloat16)
        wt = ops.transpose_w_bf16(torch.from_numpy(w).to(cuda))
        layout = ops.W_SCD
    else:
        xt, wt, layout = torch.from_numpy(x).to(cuda), torch.from_numpy(w).to(cuda), ops.W_SDC
    return z, xt, wt, torch.from_numpy(bias).to(cuda), torch.from_numpy(lt).to(cuda), layout


def _lse(z):
    m = z.max(-1, keepdims=True)
    return (m + np.log(np.exp(z - m).sum(-1, keepdims=True)))[..., 0]


@pytest.mark.parametrize(
    "S,B,D,C,bf16",
    [(3, 300, 512, 200, True),    # tcgen05, row complete in one tile: fused softmax / log_softmax
     (2, 129, 256, 64, True), (2, 256, 320, 256, True),
     (2, 256, 256, 1000, True),   # tcgen05, 4 column tiles: (max, sum-exp) partials merged
     (4, 96, 84, 10, False),      # CUDA-core contraction + row pass (LeNet-5 head)
     (3, 200, 768, 2, True)],     # narrow bf16 head (BERT classifier): regular dispatch + row pass
)
def test_softmax_epilogue_and_row_lse(cuda, S, B, D, C, bf16):
    """fb200_last_layer_logits_ex: softmax / log_softmax of ClassificationProbOutputLayer fused behind the contraction
    (prob_output_layer/classification.py:25-28, 54-69) and the per-row logsumexp; then log_prob from S*B gathers."""
    from fortuna_b200 import ops

    z, xt, wt, bias, lt, layout = _ex_inputs(cuda, S, B, D, C, seed=S + C, bf16=bf16)
    tol = 2e-3 if bf16 else 1e-5
    atol = tol * np.sqrt(D) / 8 if bf16 else 1e-4
    lse_o = _lse(z.astype(np.float64))
    out, lse = ops.last_layer_logits_ex(xt, wt, bias, lt, w_layout=layout, want_lse=True)
    np.testing.assert_allclose(out.cpu().numpy(), z, rtol=tol, atol=atol)
    np.testing.assert_allclose(lse.cpu().numpy(), lse_o, rtol=tol, atol=atol)
    # the statistics are exact for the logits that were actually produced
    got_z = out.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(lse.cpu().numpy(), _lse(got_z), rtol=1e-5, atol=1e-5)
    if C <= 256:
        sm = ops.last_layer_logits_ex(xt, wt, bias, lt, w_layout=layout, epilogue=ops.EPI_SOFTMAX).cpu().numpy()
        np.testing.assert_allclose(sm, np.exp(got_z - _lse(got_z)[..., None]), rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(sm.sum(-1), 1.0, rtol=1e-5)
        lsm, lse2 = ops.last_layer_logits_ex(xt, wt, bias, lt, w_layout=layout, epilogue=ops.EPI_LOG_SOFTMAX, want_lse=True)
        np.testing.assert_allclose(lsm.cpu().numpy(), got_z - _lse(got_z)[..., None], rtol=1e-5, atol=2e-5)
        np.testing.assert_allclose(lse2.cpu().numpy(), lse.cpu().numpy(), rtol=1e-6, atol=1e-6)
    # predictive.log_prob / ensemble_log_prob from the logits + lse (gathers) == the single-pass reduction kernel
    t = np.random.default_rng(1).integers(0, C, B).astype(np.int32)
    tt = torch.from_numpy(t).to(cuda)
    lp, ens = ops.log_prob_from_lse(out, lse, tt, want_ensemble=True)
    ref = ops.bma_reduce_cls(out, ("log_prob", "ensemble_log_prob"), targets=tt)
    np.testing.assert_allclose(lp.cpu().numpy(), ref["log_prob"].cpu().numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(ens.cpu().numpy(), ref["ensemble_log_prob"].cpu().numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(lp.cpu().numpy(), P.cls_log_prob(out.cpu().numpy(), t), rtol=1e-5, atol=2e-6)


def test_softmax_epilogue_bf16_out_and_gather(cuda):
    from fortuna_b200 import ops

    S, B, D, C = 4, 512, 512, 128
    z, xt, wt, bias, lt, layout = _ex_inputs(cuda, 6, B, D, C, seed=11)
    idx = torch.tensor([5, 0, 0, 3], dtype=torch.int32, device=cuda)
    sm = ops.last_layer_logits_ex(xt, wt, bias, lt, out_dtype=torch.bfloat16, w_layout=layout, epilogue=ops.EPI_SOFTMAX,
                                  sample_idx=idx).float().cpu().numpy()
    want = P.softmax(z[idx.cpu().numpy()])
    np.testing.assert_allclose(sm, want, rtol=2e-2, atol=2e-3)
    # wider than one tile: a normalised epilogue cannot be row-complete in TMEM, so the contraction is followed by one
    # in-place row pass - same result
    z2, xt2, wt2, b2, lt2, lay2 = _ex_inputs(cuda, 1, 256, 256, 1000, seed=1)
    sm2 = ops.last_layer_logits_ex(xt2, wt2, b2, lt2, w_layout=lay2, epilogue=ops.EPI_SOFTMAX).cpu().numpy()
    np.testing.assert_allclose(sm2, P.softmax(z2), rtol=2e-2, atol=1e-5)
    np.testing.assert_allclose(sm2.sum(-1), 1.0, rtol=1e-5)
