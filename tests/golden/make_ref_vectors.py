#!/usr/bin/env python
"""Writes tests/golden/ref_vectors.npz: seeded inputs and the outputs of the REFERENCE'S OWN SOURCE FILES for them.

The reference (/root/reference/fortuna) is pure Python over jax / flax / optax, none of which is installed here.  This
script imports the reference's hot-path modules from /root/reference as they are and executes them over the numpy
stand-in of tests/golden/jax_standin.py (see its header for exactly what that pins: the reference's control flow,
formulas, key schedule and dtype promotions - everything written in its .py files - and what it does not: the numerics
of XLA primitives).  What the reference's model stack would provide around the path is replaced by the smallest
possible doubles, all defined below:

* ``ModelManager.apply`` = features @ kernel + bias (the last Dense layer; the backbone is out of scope),
* the output calibrator = the temperature scaler's one line (output_calibrator/classification.py:16-17,
  output_calibrator/regression.py:16-19),
* the posterior state repository = a list of stored samples (SGMCMCPosterior.sample itself IS the reference's code).

/root/reference does not travel to the GPU box, so the vectors are committed; tests/test_ref_vectors.py checks the
oracle against them on CPU and tests/test_gpu_ref_vectors.py the CUDA path.

Run from the repo root:  python tests/golden/make_ref_vectors.py [--out other.npz]
(tests/test_ref_vectors.py::test_committed_vectors_are_what_the_reference_source_produces re-runs it wherever
/root/reference is present and compares every array with the committed file.)
"""
import os
import sys
import warnings
from collections import namedtuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore", category=SyntaxWarning)

from tests.golden import jax_standin as J  # noqa: E402

imp = J.load_reference()
jnp = sys.modules["jax.numpy"]
optax = sys.modules["optax"]

from oracle import prng as OP  # noqa: E402  (only PRNGKey, to make key arrays)
from tests import trees  # noqa: E402

OUT = {}
A = np.asarray


# Loader doubles (fortuna/data/loader/base.py and utils.py:169-174 restated: iteration over batches, sizes, the
# "concatenate" test grid of conformal_set, ConcatenatedLoader's chained iteration).  They are put into the stub module
# BEFORE the reference's predictive modules are imported, so those modules bind these names.
class InputsLoaderDouble:
    def __init__(self, batches):
        self.batches = list(batches)

    def __iter__(self):
        return iter(self.batches)

    @property
    def size(self):
        return sum(int(b.shape[0]) for b in self.batches)


class DataLoaderDouble:
    def __init__(self, batches, num_unique_labels=None):
        self.batches, self.num_unique_labels = batches, num_unique_labels

    def __iter__(self):
        return iter(self.batches() if callable(self.batches) else self.batches)

    @property
    def size(self):
        return sum(int(b[0].shape[0]) for b in self)

    @classmethod
    def from_inputs_loaders(cls, inputs_loaders, targets, how="interpose"):
        assert how == "concatenate"

        def inner():
            for loader, target in zip(inputs_loaders, targets):
                for inputs in loader:
                    yield inputs, J._j(target * np.ones(inputs.shape[0], dtype="int32"))

        return cls(inner)


class ConcatenatedLoaderDouble:
    def __init__(self, loaders):
        self._loaders = loaders

    def __iter__(self):
        for loader in self._loaders:
            yield from loader


_loader_mod = imp("fortuna.data.loader")
_loader_mod.DataLoader, _loader_mod.InputsLoader = DataLoaderDouble, InputsLoaderDouble
_loader_mod.ConcatenatedLoader = ConcatenatedLoaderDouble


def put(name, value):
    v = A(value)
    assert v.dtype != object, name
    OUT[name] = v


# --------------------------------------------------------------------------------------------------------------------
# doubles for the parts of the model stack that are not on the path
# --------------------------------------------------------------------------------------------------------------------
class LastLayerModelManager:
    def apply(self, params, inputs, mutable=None, **kw):
        return jnp.dot(inputs, params["kernel"]) + params["bias"]


class ClsTemperature:
    output_calibrator = object()

    def apply(self, params, outputs, mutable=None, **kw):
        return outputs * jnp.exp(-params["log_temp"])


class RegTemperature:
    output_calibrator = object()

    def apply(self, params, outputs, mutable=None, **kw):
        mean, log_var = jnp.split(outputs, 2, axis=-1)
        return jnp.concatenate((mean, log_var + params["log_temp"]), axis=-1)


State = namedtuple("State", "params mutable calib_params calib_mutable")


class Repo:
    def __init__(self, states):
        self.states = states

    def get(self, i=0, **kw):
        return self.states[int(i)]


class NS:
    pass


def make_predictive(kind, W, b, log_temp):
    post_mod = imp("fortuna.prob_model.posterior.sgmcmc.sgmcmc_posterior")
    if kind == "cls":
        lik_cls = imp("fortuna.likelihood.classification").ClassificationLikelihood
        pol = imp("fortuna.prob_output_layer.classification").ClassificationProbOutputLayer()
        pred_cls = imp("fortuna.prob_model.predictive.classification").ClassificationPredictive
        calib = ClsTemperature()
    else:
        lik_cls = imp("fortuna.likelihood.regression").RegressionLikelihood
        pol = imp("fortuna.prob_output_layer.regression").RegressionProbOutputLayer()
        pred_cls = imp("fortuna.prob_model.predictive.regression").RegressionPredictive
        calib = RegTemperature()
    n = W.shape[0]
    cal = None if log_temp is None else {"output_calibrator": {"log_temp": J._j(np.array([log_temp], np.float32))}}
    states = [State({"kernel": J._j(W[i]), "bias": J._j(b[i])}, None, cal, None) for i in range(n)]
    lik = lik_cls(LastLayerModelManager(), pol, calib if log_temp is not None else None)
    p = object.__new__(post_mod.SGMCMCPosterior)
    p.state = Repo(states)
    p.posterior_approximator = NS()
    p.posterior_approximator.n_samples = n
    p.joint = NS()
    p.joint.likelihood = lik
    return pred_cls(p)


def batches(x, sizes):
    out, o = [], 0
    for s in sizes:
        out.append(J._j(x[o : o + s]))
        o += s
    assert o == len(x)
    return out


# --------------------------------------------------------------------------------------------------------------------
# 1. predictive, classification (prob_model/predictive/base.py + classification.py, likelihood/, prob_output_layer/)
# --------------------------------------------------------------------------------------------------------------------
def gen_cls_predictive():
    for tag, (n, D, C, B, S, log_temp, seed) in {
        "a": (8, 20, 11, 50, 9, None, 1),
        "b": (6, 16, 3, 40, 30, 0.3, 2),     # S > n: draws repeat, as in the reference
        "c": (12, 32, 100, 24, 7, -0.2, 3),
        "d": (8, 16, 1000, 36, 16, 0.1, 4),  # the headline width: C = 1000 (tiled single-pass reduction, select kernel)
    }.items():
        r = np.random.default_rng(seed)
        W = (3 * r.standard_normal((n, D, C)) / np.sqrt(D)).astype(np.float32)
        b = r.standard_normal((n, C)).astype(np.float32)
        x = r.standard_normal((B, D)).astype(np.float32)
        t = r.integers(0, C, B).astype(np.int32)
        key = OP.PRNGKey(100 + seed)
        sizes = [B // 2 + 3, B - (B // 2 + 3)]
        pred = make_predictive("cls", W, b, log_temp)
        loader = batches(x, sizes)
        dloader = list(zip(batches(x, sizes), batches(t, sizes)))
        kw = dict(n_posterior_samples=S, rng=J._j(key), distribute=False)
        pre = f"cls_{tag}/"
        for k, v in dict(W=W, b=b, x=x, t=t, key=key, sizes=np.array(sizes), S=np.int32(S),
                         log_temp=np.float32(np.nan if log_temp is None else log_temp)).items():
            put(pre + k, v)
        for m in ("mean", "mode", "aleatoric_variance", "epistemic_variance", "variance", "std", "entropy",
                  "aleatoric_entropy", "epistemic_entropy"):
            put(pre + m, getattr(pred, m)(loader, **kw))
        put(pre + "log_prob", pred.log_prob(dloader, **kw))
        put(pre + "ensemble_log_prob", pred.ensemble_log_prob(dloader, **kw))
        if C < 1000:  # [S, B, C]: left out of the wide case to keep the fixture small
            put(pre + "calibrated_outputs", pred.sample_calibrated_outputs(loader, n_output_samples=S, rng=J._j(key),
                                                                           distribute=False))
        # target samples: Predictive.sample cannot take an explicit rng (base.py:355, see the regression case below)
        RNG = imp("fortuna.utils.random").RandomNumberGenerator
        put(pre + "sample_key", RNG(seed).get())
        pred.rng = RNG(seed)
        put(pre + "sample", pred.sample([J._j(x)], n_target_samples=6, distribute=False))
        cs = pred.credible_set(loader, n_posterior_samples=S, error=0.1, rng=J._j(key), distribute=False)
        sizes_cs = np.array([len(s) for s in cs], np.int32)
        padded = np.full((B, C), -1, np.int32)
        for i, s in enumerate(cs):
            padded[i, : len(s)] = s
        put(pre + "credible_set", padded)
        put(pre + "credible_set_sizes", sizes_cs)


# --------------------------------------------------------------------------------------------------------------------
# 2. predictive, regression
# --------------------------------------------------------------------------------------------------------------------
def gen_reg_predictive():
    for tag, (n, D, d, B, S, T, log_temp, seed) in {
        "a": (7, 12, 1, 30, 8, 5, None, 11),
        "b": (5, 10, 3, 26, 6, 4, 0.25, 12),
    }.items():
        r = np.random.default_rng(seed)
        W = (r.standard_normal((n, D, 2 * d)) / np.sqrt(D)).astype(np.float32)
        b = (0.3 * r.standard_normal((n, 2 * d))).astype(np.float32)
        x = r.standard_normal((B, D)).astype(np.float32)
        y = r.standard_normal((B, d)).astype(np.float32)
        key = OP.PRNGKey(200 + seed)
        sizes = [B // 2 + 1, B - (B // 2 + 1)]
        pred = make_predictive("reg", W, b, log_temp)
        loader = batches(x, sizes)
        dloader = list(zip(batches(x, sizes), batches(y, sizes)))
        kw = dict(n_posterior_samples=S, rng=J._j(key), distribute=False)
        pre = f"reg_{tag}/"
        for k, v in dict(W=W, b=b, x=x, y=y, key=key, sizes=np.array(sizes), S=np.int32(S), T=np.int32(T),
                         log_temp=np.float32(np.nan if log_temp is None else log_temp)).items():
            put(pre + k, v)
        for m in ("mean", "mode", "aleatoric_variance", "epistemic_variance", "variance", "std"):
            put(pre + m, getattr(pred, m)(loader, **kw))
        put(pre + "log_prob", pred.log_prob(dloader, **kw))
        put(pre + "ensemble_log_prob", pred.ensemble_log_prob(dloader, **kw))
        # Monte-Carlo entropies: one input batch (the reference re-uses rng per batch; the target draws depend on the
        # batch shape, so a single batch keeps the fixture independent of the batching)
        one = [J._j(x)]
        for m in ("entropy", "aleatoric_entropy", "epistemic_entropy"):
            put(pre + m, getattr(pred, m)(one, n_posterior_samples=S, n_target_samples=T, rng=J._j(key),
                                          distribute=False))
        # Predictive.sample tests ``if not rng`` (base.py:355), which raises for a key array (in jax as here): an explicit
        # rng cannot be passed, the key comes from the object's RandomNumberGenerator.  The fixture records that key.
        RNG = imp("fortuna.utils.random").RandomNumberGenerator
        put(pre + "sample_key", RNG(seed).get())
        pred.rng = RNG(seed)
        put(pre + "sample", pred.sample(one, n_target_samples=T, distribute=False))
        pred.rng = RNG(seed)
        put(pre + "quantile", pred.quantile([0.05, 0.5, 0.95], one, n_target_samples=T, distribute=False))


# --------------------------------------------------------------------------------------------------------------------
# 3. samplers: sghmc_integrator / cyclical_sgld_integrator / preconditioners / schedules / gradient clipping
# --------------------------------------------------------------------------------------------------------------------
def grad_of(p_np):
    """A deterministic 'gradient' both sides can form: g = 0.5 p + 0.01 (float32)."""
    return {k: grad_of(v) for k, v in p_np.items()} if isinstance(p_np, dict) else (p_np * 0.5 + 0.01).astype(np.float32)


def flat_leaves(tree):
    return [A(x) for x in J.tree_leaves(tree)]


def gen_samplers():
    sgi = imp("fortuna.prob_model.posterior.sgmcmc.sghmc.sghmc_integrator")
    pcm = imp("fortuna.prob_model.posterior.sgmcmc.sgmcmc_preconditioner")
    sch = imp("fortuna.prob_model.posterior.sgmcmc.sgmcmc_step_schedule")
    cyc = imp("fortuna.prob_model.posterior.sgmcmc.cyclical_sgld.cyclical_sgld_integrator")
    tr = imp("fortuna.utils.training")
    # the two-moons MLP (1 082 parameters) with every step stored; the ragged tree (leaves of 1 .. 12 288 elements, odd
    # sizes) with the last step stored
    n_steps = 6
    put("sampler/n_steps", np.int32(n_steps))

    def run(tag, make_tx, state_of):
        for tree_name, params, stored in (("mlp", trees.mlp_two_moons_tree(0), range(n_steps)),
                                          ("ragged", trees.ragged_tree(0), (n_steps - 1,))):
            jp = J.tree_map(J._j, params)
            tx = make_tx()
            st, p = tx.init(jp), jp
            for step in range(n_steps):
                g = J.tree_map(J._j, grad_of(J.tree_map(A, p)))
                upd, st = tx.update(g, st)
                p = optax.apply_updates(p, upd)
                s = state_of(st)
                pre = f"sampler/{tag}/{tree_name}/step{step}/"
                put(pre + "key", s.rng_key)
                put(pre + "count", s.count)
                if step not in stored:
                    continue
                put(pre + "params", np.concatenate([x.ravel() for x in flat_leaves(p)]))
                put(pre + "momentum", np.concatenate([x.ravel() for x in flat_leaves(s.momentum)]))
                v = flat_leaves(s.preconditioner_state)
                if v:
                    put(pre + "v", np.concatenate([x.ravel() for x in v]))

    def precond_of(name):
        return pcm.rmsprop_preconditioner() if name == "rmsprop" else pcm.identity_preconditioner()

    for precond in ("identity", "rmsprop"):
        for resample in (None, 3):
            run(f"sghmc_{precond}_{resample}",
                lambda: sgi.sghmc_integrator(0.05, resample, J._j(OP.PRNGKey(7)), sch.cosine_schedule(1e-3, 10),
                                             precond_of(precond)),
                lambda st: st)
        # cycle of 4 at exploration ratio 0.5: two exploration steps, then two sampling steps, ...
        run(f"csgld_{precond}",
            lambda: cyc.cyclical_sgld_integrator(J._j(OP.PRNGKey(9)), 1e-3, 4, 0.5, precond_of(precond)),
            lambda st: st.sgld_state)

    # step schedules over a range of steps (int32 counts, as the integrator passes them)
    counts = np.arange(0, 25, dtype=np.int32)
    scheds = {
        "constant": sch.constant_schedule(2e-3),
        "cosine": sch.cosine_schedule(1e-2, 10),
        "polynomial": sch.polynomial_schedule(a=0.5, b=2.0, gamma=0.55),
        "cosine_burnin": sch.constant_schedule_with_cosine_burnin(1e-2, 1e-3, 8),
        "cyclical": sch.cyclical_cosine_schedule_with_const_burnin(1e-2, 5, 6),
    }
    put("schedule/counts", counts)
    for name, fn in scheds.items():
        put(f"schedule/{name}", np.array([A(fn(J._j(c))) for c in counts], dtype=np.float32))

    # gradient clipping by global norm (utils/training.py:14-20)
    g = grad_of(trees.ragged_tree(0))
    for i, mx in enumerate((0.5, 1e3)):
        clipped = tr.clip_grandients_by_norm(J.tree_map(J._j, g), mx)
        put(f"clip/{i}/max_norm", np.float32(mx))
        put(f"clip/{i}/grad", np.concatenate([x.ravel() for x in flat_leaves(clipped)]))


# --------------------------------------------------------------------------------------------------------------------
# 4. scoring: APS / simple scores, conformal sets, ECE / MCE / accuracy / Brier
# --------------------------------------------------------------------------------------------------------------------
def gen_scoring():
    aps = imp("fortuna.conformal.classification.adaptive_prediction")
    simple = imp("fortuna.conformal.classification.simple_prediction")
    met = imp("fortuna.metric.classification")
    sm = sys.modules["jax.nn"].softmax
    for tag, (nv, nt, C, scale, seed) in {"a": (200, 64, 37, 4.0, 0), "b": (500, 100, 10, 6.0, 1),
                                          "c": (64, 20, 1000, 8.0, 2)}.items():
        r = np.random.default_rng(seed)
        vp = A(sm(J._j((scale * r.standard_normal((nv, C))).astype(np.float32))))
        tp = A(sm(J._j((scale * r.standard_normal((nt, C))).astype(np.float32))))
        if tag == "b":  # exact ties and a degenerate row
            tp[0] = np.float32(0.1)
            tp[1, :] = 0
            tp[1, 3] = 1
            vp[0] = vp[1]
        # targets drawn mostly from the likely classes so that the sets are small
        vt = np.array([r.choice(C, p=p.astype(np.float64) / p.astype(np.float64).sum()) for p in vp]).astype(np.int32)
        pre = f"score_{tag}/"
        put(pre + "val_probs", vp)
        put(pre + "test_probs", tp)
        put(pre + "val_targets", vt)
        jvp, jtp, jvt = J._j(vp), J._j(tp), J._j(vt)
        for kind, cl in (("aps", aps.AdaptivePredictionConformalClassifier()),
                         ("simple", simple.SimplePredictionConformalClassifier())):
            vs, ts = cl.get_scores(jvp, jvt, jtp)
            put(pre + kind + "_val_scores", vs)
            put(pre + kind + "_test_scores", ts)
            for e in (0.05, 0.2) + ((0.0, 1.0, 0.999999) if tag == "a" else ()):  # the extremes: every class / none
                sets = cl.conformal_set(jvp, jvt, jtp, error=e)
                mask = np.zeros((nt, C), np.uint8)
                for i, s in enumerate(sets):
                    mask[i, s] = 1
                put(pre + f"{kind}_sets_{e}", mask)
        if tag == "a":  # CV+ variants: two folds (conformal/classification/base.py:92-132)
            h = nv // 2
            for kind, cl in (("aps", aps.CVPlusAdaptivePredictionConformalClassifier()),
                             ("simple", simple.CVPlusSimplePredictionConformalClassifier())):
                sets = cl.conformal_set([jvp[:h], jvp[h:]], [jvt[:h], jvt[h:]], [jtp[: nt // 2], jtp[nt // 2 :]], error=0.1)
                mask = np.zeros((nt, C), np.uint8)
                for i, s_ in enumerate(sets):
                    mask[i, s_] = 1
                put(pre + f"cvplus_{kind}_sets_0.1", mask)
        preds = vp.argmax(-1).astype(np.int32)
        jpreds = J._j(preds)
        counts, confs, accs = met.compute_counts_confs_accs(jpreds, jvp, jvt)
        put(pre + "ece_counts", counts)
        put(pre + "ece_confs", confs)
        put(pre + "ece_accs", accs)
        put(pre + "ece", met.expected_calibration_error(jpreds, jvp, jvt))
        put(pre + "mce", met.maximum_calibration_error(jpreds, jvp, jvt))
        put(pre + "accuracy", met.accuracy(jpreds, jvt))
        put(pre + "brier", met.brier_score(jvp, jvt))


# --------------------------------------------------------------------------------------------------------------------
# 5. diagnostics: KSD / ESS (sgmcmc_diagnostic.py)
# --------------------------------------------------------------------------------------------------------------------
def gen_diagnostics():
    dg = imp("fortuna.prob_model.posterior.sgmcmc.sgmcmc_diagnostic")
    r = np.random.default_rng(4)
    S, shapes = 12, {"a": (5, 3), "b": (7,)}
    samples, grads = [], []
    prev = {k: r.standard_normal(s).astype(np.float32) for k, s in shapes.items()}
    for _ in range(S):  # an AR(1) chain so that the autocorrelations are not all noise
        prev = {k: (0.7 * v + 0.5 * r.standard_normal(v.shape)).astype(np.float32) for k, v in prev.items()}
        samples.append(prev)
        grads.append({k: (-v + 0.1 * r.standard_normal(v.shape)).astype(np.float32) for k, v in prev.items()})
    put("diag/samples", np.stack([np.concatenate([s[k].ravel() for k in sorted(s)]) for s in samples]))
    put("diag/grads", np.stack([np.concatenate([s[k].ravel() for k in sorted(s)]) for s in grads]))
    js = [J.tree_map(J._j, s) for s in samples]
    jg = [J.tree_map(J._j, s) for s in grads]
    put("diag/ksd", dg.kernel_stein_discrepancy_imq(js, jg))
    put("diag/ksd_c2_b075", dg.kernel_stein_discrepancy_imq(js, jg, c=2.0, beta=-0.75))
    ess = dg.effective_sample_size(js)
    put("diag/ess", np.concatenate([A(ess[k]).ravel() for k in sorted(ess)]))


# --------------------------------------------------------------------------------------------------------------------
# 6. multivalid scorers (conformal/multivalid/): iterative (Multicalibrator, BatchMVP, top-label, binary) and one-shot
# --------------------------------------------------------------------------------------------------------------------
def _mv_data(n, seed=0):
    r = np.random.default_rng(seed)
    x = r.random(n).astype(np.float32)
    scores = np.clip(0.7 * x + 0.3 * r.random(n), 0, 1).astype(np.float32)
    groups = np.stack([x < 0.5, r.random(n) < 0.3], axis=1)
    return x, scores, groups


def _put_run(pre, status, patches):
    put(pre + "losses", np.array([float(A(v)) for v in status["losses"]], np.float32))
    put(pre + "n_rounds", np.int32(status["n_rounds"]))
    put(pre + "converged", np.int32(bool(status["converged"])))
    # patches: (tau, g, v, c, patch) per round
    put(pre + "patches", np.array([[float(A(t)) for t in p] for p in patches], np.float64).reshape(len(patches), 5))
    print(f"  {pre} rounds kept: {len(patches)}")


def gen_multivalid():
    logging_off()
    mc = imp("fortuna.conformal.multivalid.iterative.multicalibrator")
    bm = imp("fortuna.conformal.multivalid.iterative.regression.batch_mvp")
    tlm = imp("fortuna.conformal.multivalid.iterative.classification.top_label_multicalibrator")
    bcm = imp("fortuna.conformal.multivalid.iterative.classification.binary_multicalibrator")
    osm = imp("fortuna.conformal.multivalid.one_shot.multicalibrator")
    osb = imp("fortuna.conformal.multivalid.one_shot.classification.binary_multicalibrator")
    ost = imp("fortuna.conformal.multivalid.one_shot.classification.top_label_multicalibrator")

    x, scores, groups = _mv_data(500)
    tx, ts, tg = _mv_data(120, seed=5)
    v0 = np.clip(x + 0.2, 0, 1).astype(np.float32)
    for k, v in dict(x=x, scores=scores, groups=groups, tx=tx, ts=ts, tg=tg, v0=v0).items():
        put("mv/data/" + k, v)
    for tag, kw in {"add": dict(bucket_types=("<=", ">="), patch_type="additive"),
                    "mul": dict(bucket_types=("=", ">=", "<="), patch_type="multiplicative")}.items():
        m = mc.Multicalibrator(seed=1)
        tv, status = m.calibrate(scores=J._j(scores), groups=J._j(groups), values=J._j(v0), test_groups=J._j(tg),
                                 test_values=J._j(tx), n_rounds=6, n_buckets=8, eta=0.5, **kw)
        _put_run(f"mv/multicalibrator_{tag}/", status, m.patches)
        put(f"mv/multicalibrator_{tag}/test_values", tv)
        put(f"mv/multicalibrator_{tag}/calibration_error", m.calibration_error(J._j(ts), J._j(tg), tv))
    b = bm.BatchMVPConformalClassifier(seed=2)
    status = b.calibrate(scores=J._j(scores), groups=J._j(groups), n_rounds=5, n_buckets=10, eta=0.5, coverage=0.9)
    _put_run("mv/batchmvp/", status, b.patches)
    th = b.apply_patches(J._j(tg))
    put("mv/batchmvp/test_thresholds", th)
    put("mv/batchmvp/calibration_error", b.calibration_error(J._j(ts), J._j(tg), th))

    r = np.random.default_rng(4)
    n, C = 700, 4
    logits = r.standard_normal((n, C)).astype(np.float32)
    probs = (np.exp(logits) / np.exp(logits).sum(1, keepdims=True)).astype(np.float32)
    targets = np.array([r.choice(C, p=q / q.sum()) for q in probs.astype(np.float64)], dtype=np.int32)
    probs = (np.exp(3 * logits) / np.exp(3 * logits).sum(1, keepdims=True)).astype(np.float32)  # over-confident model
    cg = r.random((n, 2)) < 0.5
    for k, v in dict(probs=probs, targets=targets, groups=cg).items():
        put("mv/cls/" + k, v)
    t = tlm.TopLabelMulticalibrator(n_classes=C, seed=0)
    status = t.calibrate(targets=J._j(targets), groups=J._j(cg), probs=J._j(probs), n_rounds=4, n_buckets=5, eta=1.0)
    _put_run("mv/top_label/", status, t.patches)
    put("mv/top_label/applied", t.apply_patches(J._j(cg), J._j(probs)))
    y = (r.random(n) < probs[:, 0] ** 2).astype(np.int32)  # the model over-predicts the positive class
    put("mv/cls/binary_targets", y)
    bc = bcm.BinaryClassificationMulticalibrator(seed=0)
    status = bc.calibrate(targets=J._j(y), groups=J._j(cg), probs=J._j(probs[:, 0].copy()), n_rounds=4, n_buckets=6, eta=0.5)
    _put_run("mv/binary/", status, bc.patches)

    # one-shot: repeated values, several data points per unique value
    xr = np.round(x, 2).astype(np.float32)
    txr = np.round(tx, 2).astype(np.float32)
    put("mv/one_shot/values", xr)
    put("mv/one_shot/test_values", txr)
    o = osm.OneShotMulticalibrator()
    tv = o.calibrate(scores=J._j(scores), values=J._j(xr), test_values=J._j(txr), n_buckets=10)
    put("mv/one_shot/patches", o.patches)
    put("mv/one_shot/applied", tv)
    yb = (np.random.default_rng(1).random(len(xr)) < xr).astype(np.int32)
    put("mv/one_shot/binary_targets", yb)
    ob = osb.OneShotBinaryClassificationMulticalibrator()
    ob.calibrate(targets=J._j(yb), probs=J._j(xr), n_buckets=10)
    put("mv/one_shot/binary_patches", ob.patches)
    pr = np.round(np.random.default_rng(2).dirichlet(np.ones(3), 300), 2).astype(np.float32)
    tt = np.random.default_rng(3).integers(0, 3, 300).astype(np.int32)
    put("mv/one_shot/tl_probs", pr)
    put("mv/one_shot/tl_targets", tt)
    ot = ost.OneShotTopLabelMulticalibrator(n_classes=3)
    out = ot.calibrate(targets=J._j(tt), probs=J._j(pr), test_probs=J._j(pr), n_buckets=10)
    put("mv/one_shot/tl_patches", ot.patches)
    put("mv/one_shot/tl_applied", out)


# --------------------------------------------------------------------------------------------------------------------
# 7. MaxCoverageFixedPrecision binary calibrator (conformal/classification/maxcovfixprec_binary_classfication.py)
# --------------------------------------------------------------------------------------------------------------------
def gen_maxcov():
    logging_off()
    mod = imp("fortuna.conformal.classification.maxcovfixprec_binary_classfication")
    r = np.random.default_rng(6)
    n = 800
    q = r.random(n).astype(np.float32)
    y = (r.random(n) < q ** 1.3).astype(np.int32)
    pt = r.random(200).astype(np.float32)
    for k, v in dict(probs=q, targets=y, test_probs=pt).items():
        put("maxcov/" + k, v)
    for i, (tp, tn, margin) in enumerate(((0.8, 0.75, 0.0), (0.9, 0.9, 0.02))):
        cal = mod.MaxCoverageFixedPrecisionBinaryClassificationCalibrator()
        out = cal.calibrate(targets=J._j(y), probs=J._j(q), true_positive_precision_threshold=tp,
                            true_negative_precision_threshold=tn, test_probs=J._j(pt), n_taus=60, margin=margin)
        put(f"maxcov/{i}/thresholds", np.array([tp, tn, margin], np.float64))
        put(f"maxcov/{i}/tau_pos", cal.patches["tau_pos"])
        put(f"maxcov/{i}/tau_neg", cal.patches["tau_neg"])
        put(f"maxcov/{i}/applied", out)
        put(f"maxcov/{i}/tp_precision", cal.true_positive_precision(out, J._j(y[:200]), tp))
        put(f"maxcov/{i}/tn_precision", cal.true_negative_precision(out, J._j(y[:200]), tn))


# --------------------------------------------------------------------------------------------------------------------
# 8. calibration losses (loss/classification/focal_loss.py, loss/regression/scaled_mse.py)
# --------------------------------------------------------------------------------------------------------------------
def gen_losses():
    fl = imp("fortuna.loss.classification.focal_loss").focal_loss_fn
    sm = imp("fortuna.loss.regression.scaled_mse").scaled_mse_fn
    r = np.random.default_rng(8)
    z = (3 * r.standard_normal((300, 7))).astype(np.float32)
    t = r.integers(0, 7, 300).astype(np.int32)
    put("loss/logits", z)
    put("loss/targets", t)
    for g in (2.0, 0.5):
        put(f"loss/focal_gamma{g}", fl(J._j(z), J._j(t), gamma=g))
    o = r.standard_normal((200, 6)).astype(np.float32)
    y = r.standard_normal((200, 3)).astype(np.float32)
    put("loss/outputs", o)
    put("loss/y", y)
    put("loss/scaled_mse", sm(J._j(o), J._j(y)))


# --------------------------------------------------------------------------------------------------------------------
# 9. conformal Bayes sets (prob_model/predictive/classification.py:485-626)
# --------------------------------------------------------------------------------------------------------------------
def gen_conformal_bayes():
    n, D, C, n_train, n_test, S, seed = 9, 12, 5, 60, 25, 8, 21
    r = np.random.default_rng(seed)
    # stored samples scattered around one confident model, training labels from that model: the sets are selective
    W0 = (5 * r.standard_normal((D, C)) / np.sqrt(D)).astype(np.float32)
    W = (W0[None] + 0.6 * r.standard_normal((n, D, C)) / np.sqrt(D)).astype(np.float32)
    b = (0.3 * r.standard_normal((n, C))).astype(np.float32)
    xtr = r.standard_normal((n_train, D)).astype(np.float32)
    ytr = (xtr @ W0).argmax(1).astype(np.int32)
    flip = r.random(n_train) < 0.15
    ytr[flip] = r.integers(0, C, int(flip.sum()))
    xte = r.standard_normal((n_test, D)).astype(np.float32)
    key = OP.PRNGKey(300 + seed)
    pred = make_predictive("cls", W, b, None)
    train = DataLoaderDouble([(J._j(xtr[:32]), J._j(ytr[:32])), (J._j(xtr[32:]), J._j(ytr[32:]))], num_unique_labels=C)
    test = InputsLoaderDouble([J._j(xte[:16]), J._j(xte[16:])])
    for k, v in dict(W=W, b=b, x_train=xtr, y_train=ytr, x_test=xte, key=key, S=np.int32(S)).items():
        put("cb/" + k, v)
    for e in (0.1, 0.4):
        sets, ess = pred.conformal_set(train, test, n_posterior_samples=S, error=e, rng=J._j(key), distribute=False,
                                       return_ess=True)
        mask = np.zeros((n_test, C), np.uint8)
        for i, s_ in enumerate(sets):
            mask[i, s_] = 1
        put(f"cb/region_{e}", mask)
    put("cb/ess", np.asarray(ess).reshape(n_test, C))


# --------------------------------------------------------------------------------------------------------------------
# 10. conformal regression (conformal/regression/quantile.py, onedim_uncertainty.py, base.py)
# --------------------------------------------------------------------------------------------------------------------
def gen_conformal_regression():
    qm = imp("fortuna.conformal.regression.quantile")
    om = imp("fortuna.conformal.regression.onedim_uncertainty")
    r = np.random.default_rng(13)
    nv, nt = 137, 21
    vy = r.standard_normal((nv, 1)).astype(np.float32)
    vlo = (vy[:, 0] - 1.0 + 0.7 * r.standard_normal(nv)).astype(np.float32)
    vhi = (vlo + 2.0 + 0.3 * r.random(nv)).astype(np.float32)
    tlo = r.standard_normal(nt).astype(np.float32)
    thi = (tlo + 2.0).astype(np.float32)
    for k, v in dict(val_targets=vy, val_lower=vlo, val_upper=vhi, test_lower=tlo, test_upper=thi).items():
        put("creg/" + k, v)
    c = qm.QuantileConformalRegressor()
    put("creg/q_scores", c.score(J._j(vlo), J._j(vhi), J._j(vy)))
    for e in (0.05, 0.2):
        put(f"creg/q_quantile_{e}", c.quantile(J._j(vlo), J._j(vhi), J._j(vy), e))
        ci = c.conformal_interval(J._j(vlo), J._j(vhi), J._j(tlo), J._j(thi), J._j(vy), e)
        put(f"creg/q_interval_{e}", ci)
        put(f"creg/q_is_in_{e}", c.is_in(J._j(np.zeros(nt, np.float32)), ci))
    vp = r.standard_normal((nv, 1)).astype(np.float32)
    vu = (0.2 + r.random((nv, 1))).astype(np.float32)
    tp = r.standard_normal((nt, 1)).astype(np.float32)
    tu = (0.2 + r.random((nt, 1))).astype(np.float32)
    for k, v in dict(val_preds=vp, val_unc=vu, test_preds=tp, test_unc=tu).items():
        put("creg/" + k, v)
    o = om.OneDimensionalUncertaintyConformalRegressor()
    put("creg/o_scores", o.score(J._j(vp), J._j(vu), J._j(vy)))
    for e in (0.05, 0.2):
        put(f"creg/o_quantile_{e}", o.quantile(J._j(vp), J._j(vu), J._j(vy), e))
        put(f"creg/o_interval_{e}", o.conformal_interval(J._j(vp), J._j(vu), J._j(tp), J._j(tu), J._j(vy), e))


# --------------------------------------------------------------------------------------------------------------------
# 11. calibration loss with ensemble outputs (predictive/base.py:205-316, likelihood/base.py:124-262)
# --------------------------------------------------------------------------------------------------------------------
def gen_calibration_loss():
    r = np.random.default_rng(31)
    S, B, C, d = 5, 40, 6, 2
    z = (2 * r.standard_normal((S, B, C))).astype(np.float32)
    t = r.integers(0, C, B).astype(np.int32)
    zr = r.standard_normal((S, B, 2 * d)).astype(np.float32)
    y = r.standard_normal((B, d)).astype(np.float32)
    for k, v in dict(logits=z, targets=t, outputs=zr, y=y, n_data=np.int32(3 * B)).items():
        put("calib/" + k, v)
    W = np.zeros((1, 3, C), np.float32)
    for kind, ens, tg, ncol in (("cls", z, t, C), ("reg", zr, y, 2 * d)):
        pred = make_predictive(kind, np.zeros((1, 3, ncol), np.float32), np.zeros((1, ncol), np.float32), 0.0)
        for lt in (0.0, 0.4, -0.3):
            cal = {"output_calibrator": {"log_temp": J._j(np.array([lt], np.float32))}}
            v = pred._batched_negative_log_joint_prob((J._j(np.zeros((B, 3), np.float32)), J._j(tg)), 3 * B,
                                                      return_aux=[], ensemble_outputs=J._j(ens), calib_params=cal,
                                                      calib_mutable=None)
            put(f"calib/{kind}_loss_{lt}", v)


# --------------------------------------------------------------------------------------------------------------------
# 12. output calibration models: S = 1 predictive maps over given outputs (output_calib_model/predictive/*.py over
#     prob_output_layer/*.py)
# --------------------------------------------------------------------------------------------------------------------
def gen_output_calib_predictive():
    cm = imp("fortuna.output_calib_model.predictive.classification")
    rm = imp("fortuna.output_calib_model.predictive.regression")
    polc = imp("fortuna.prob_output_layer.classification").ClassificationProbOutputLayer()
    polr = imp("fortuna.prob_output_layer.regression").RegressionProbOutputLayer()
    RNG = imp("fortuna.utils.random").RandomNumberGenerator

    class StateRepo:
        def __init__(self, lt):
            self.s = NS()
            self.s.params = {"output_calibrator": {"log_temp": J._j(np.array([lt], np.float32))}}
            self.s.mutable = {"output_calibrator": None}

        def get(self):
            return self.s

    r = np.random.default_rng(41)
    z = (2 * r.standard_normal((60, 7))).astype(np.float32)
    t = r.integers(0, 7, 60).astype(np.int32)
    o = r.standard_normal((50, 4)).astype(np.float32)
    y = r.standard_normal((50, 2)).astype(np.float32)
    key = OP.PRNGKey(77)
    for k, v in dict(logits=z, targets=t, outputs=o, y=y, key=key, log_temp=np.float32(0.35)).items():
        put("ocal/" + k, v)
    pc = cm.ClassificationPredictive(output_calib_manager=ClsTemperature(), prob_output_layer=polc)
    pc.state = StateRepo(0.35)
    for cal in (False, True):
        pre = f"ocal/cls_{int(cal)}/"
        for m in ("mean", "mode", "variance", "std", "entropy"):
            put(pre + m, getattr(pc, m)(J._j(z), calibrated=cal))
        put(pre + "log_prob", pc.log_prob(J._j(z), J._j(t), calibrated=cal))
        put(pre + "sample", pc.sample(9, J._j(z), rng=J._j(key), calibrated=cal))
    pr = rm.RegressionPredictive(output_calib_manager=RegTemperature(), prob_output_layer=polr)
    pr.state = StateRepo(0.35)
    for cal in (False, True):
        pre = f"ocal/reg_{int(cal)}/"
        for m in ("mean", "mode", "variance", "std"):
            put(pre + m, getattr(pr, m)(J._j(o), calibrated=cal))
        put(pre + "log_prob", pr.log_prob(J._j(o), J._j(y), calibrated=cal))
        put(pre + "sample", pr.sample(6, J._j(o), rng=J._j(key), calibrated=cal))
        put(pre + "entropy", pr.entropy(J._j(o), calibrated=cal, n_target_samples=6, rng=J._j(key)))
        put(pre + "quantile", pr.quantile([0.1, 0.9], J._j(o), n_samples=6, rng=J._j(key), calibrated=cal))


# --------------------------------------------------------------------------------------------------------------------
# 13. which training steps store a posterior sample (sghmc_callback.py, cyclical_sgld_callback.py,
#     sgmcmc_sampling_callback.py)
# --------------------------------------------------------------------------------------------------------------------
def gen_sampling_callbacks():
    sg = imp("fortuna.prob_model.posterior.sgmcmc.sghmc.sghmc_callback")
    cy = imp("fortuna.prob_model.posterior.sgmcmc.cyclical_sgld.cyclical_sgld_callback")

    class Trainer:
        def _sync_state(self, state):
            return state

    class Repo:
        def __init__(self):
            self.puts = []

        def put(self, state, i, keep):
            self.puts.append((i, state))

    def run(cb, repo, n_steps):
        for step in range(1, n_steps + 1):
            cb.training_step_end(step)  # the "state" is the 1-based step number
        return np.array(repo.puts, np.int32).reshape(-1, 2)

    cases = {
        "sghmc_0": dict(n_epochs=4, n_training_steps=25, n_samples=7, n_thinning=5, burnin_length=40),
        "sghmc_1": dict(n_epochs=3, n_training_steps=10, n_samples=30, n_thinning=1, burnin_length=0),
        "csgld_0": dict(n_epochs=5, n_training_steps=20, n_samples=9, n_thinning=2, cycle_length=20, exploration_ratio=0.25),
        "csgld_1": dict(n_epochs=2, n_training_steps=30, n_samples=6, n_thinning=3, cycle_length=15, exploration_ratio=0.6),
    }
    for tag, kw in cases.items():
        repo = Repo()
        cls = sg.SGHMCSamplingCallback if tag.startswith("sghmc") else cy.CyclicalSGLDSamplingCallback
        cb = cls(trainer=Trainer(), state_repository=repo, keep_top_n_checkpoints=1, **kw)
        put(f"callback/{tag}/stored", run(cb, repo, kw["n_epochs"] * kw["n_training_steps"]))
        put(f"callback/{tag}/config", np.array([float(v) for v in kw.values()], np.float64))
    # too few collectable samples: the constructor refuses
    for tag, cls, kw in (("sghmc", sg.SGHMCSamplingCallback, dict(n_epochs=1, n_training_steps=10, n_samples=5, n_thinning=3,
                                                                  burnin_length=4)),
                         ("csgld", cy.CyclicalSGLDSamplingCallback, dict(n_epochs=1, n_training_steps=12, n_samples=4,
                                                                         n_thinning=2, cycle_length=6, exploration_ratio=0.5))):
        try:
            cls(trainer=Trainer(), state_repository=Repo(), keep_top_n_checkpoints=1, **kw)
            refused = 0
        except ValueError:
            refused = 1
        put(f"callback/{tag}_refuses", np.int32(refused))


def logging_off():
    import logging

    logging.disable(logging.WARNING)


def main():
    gen_cls_predictive()
    gen_reg_predictive()
    gen_samplers()
    gen_scoring()
    gen_diagnostics()
    gen_multivalid()
    gen_maxcov()
    gen_losses()
    gen_conformal_bayes()
    gen_conformal_regression()
    gen_calibration_loss()
    gen_output_calib_predictive()
    gen_sampling_callbacks()
    path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **OUT)
    print(f"wrote {path}: {len(OUT)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
