"""The host-side mirrors of the reference's scoring API, exercised the way the reference's own tests do
(tests/fortuna/test_conformal_methods.py:37-100, tests/fortuna/test_metric.py) and checked value-for-value
against the oracle - same class / function names, numpy inputs like the reference tests use."""
import numpy as np
import pytest
import torch

from oracle import predictive as P
from oracle import scoring as Sc

pytestmark = pytest.mark.gpu


def test_prediction_conformal_classifier(cuda):
    from fortuna_b200.conformal.classification.simple_prediction import SimplePredictionConformalClassifier

    val_probs = np.array([[0.1, 0.9], [0.8, 0.2], [0.3, 0.7]])
    val_targets = np.array([1, 1, 0])
    test_probs = np.array([[0.5, 0.5], [0.8, 0.2]])
    conformal = SimplePredictionConformalClassifier()
    coverage_sets = conformal.conformal_set(val_probs=val_probs, test_probs=test_probs, val_targets=val_targets, error=0.05)
    assert (0 in coverage_sets[0]) and (1 in coverage_sets[0]) and (0 in coverage_sets[1])
    assert conformal.is_in(np.array([0, 0]), coverage_sets).all()


def test_simple_and_adaptive_prediction_conformal_classifiers(cuda):
    from fortuna_b200.conformal.classification.adaptive_prediction import (
        AdaptivePredictionConformalClassifier, CVPlusAdaptivePredictionConformalClassifier)
    from fortuna_b200.conformal.classification.simple_prediction import (
        CVPlusSimplePredictionConformalClassifier, SimplePredictionConformalClassifier)

    val_probs = np.array([[0.1, 0.7, 0.2], [0.5, 0.4, 0.1], [0.15, 0.6, 0.35]])
    val_targets = np.array([1, 1, 2])
    test_probs = np.array([[0.2, 0.5, 0.3], [0.6, 0.01, 0.39]])
    for split_cls, cv_cls, kind in ((SimplePredictionConformalClassifier, CVPlusSimplePredictionConformalClassifier, "simple"),
                                    (AdaptivePredictionConformalClassifier, CVPlusAdaptivePredictionConformalClassifier, "aps")):
        conformal = split_cls()
        coverage_sets = conformal.conformal_set(val_probs, val_targets, test_probs, error=0.05)
        assert len(coverage_sets) == len(test_probs)
        vs, ts = Sc.get_scores(val_probs, val_targets, test_probs, kind)
        assert coverage_sets == Sc.conformal_sets(vs, ts, 0.05)[2]
        gvs, gts = conformal.get_scores(val_probs, val_targets, test_probs)
        assert np.array_equal(gvs.cpu().numpy(), vs) and np.array_equal(gts.cpu().numpy(), ts)
        conformal = cv_cls()
        coverage_sets = conformal.conformal_set([val_probs, val_probs], [val_targets, val_targets],
                                                [test_probs, test_probs], error=0.05)
        assert len(coverage_sets) == 2 * len(test_probs)
        vs2, ts2 = np.concatenate([vs, vs]), np.concatenate([ts, ts], axis=0)
        assert coverage_sets == Sc.conformal_sets(vs2, ts2, 0.05)[2]


@pytest.mark.parametrize("error", [0.0, 0.05, 0.3, 1.0])
def test_aps_sets_larger_problem_bit_exact(cuda, error):
    from fortuna_b200.conformal.classification.adaptive_prediction import AdaptivePredictionConformalClassifier

    r = np.random.default_rng(5)
    val_probs = P.softmax((3 * r.standard_normal((777, 33))).astype(np.float32))
    val_targets = r.integers(0, 33, 777)
    test_probs = P.softmax(np.round(2 * r.standard_normal((201, 33))).astype(np.float32))  # tie-heavy
    sets = AdaptivePredictionConformalClassifier().conformal_set(val_probs, val_targets, test_probs, error=error)
    vs, ts = Sc.get_scores(val_probs, val_targets, test_probs, "aps")
    assert sets == Sc.conformal_sets(vs, ts, error)[2]


def test_quantile_conformal_regressor(cuda):
    from fortuna_b200.conformal.regression.quantile import (OneDimensionalUncertaintyConformalRegressor,
                                                            QuantileConformalRegressor)

    r = np.random.default_rng(0)
    scores = r.normal(size=100).astype(np.float32)
    for cls in (QuantileConformalRegressor, OneDimensionalUncertaintyConformalRegressor):
        quantile = cls().quantile(None, None, None, 0.05, scores=scores)
        assert scores.min() <= float(quantile) <= scores.max()
        assert float(quantile) == float(Sc.conformal_quantile(scores, 0.05))
    with pytest.raises(ValueError):
        QuantileConformalRegressor().quantile(error=1.5, scores=scores)


def test_metrics_match_oracle_and_reference_checks(cuda):
    from fortuna_b200.metric.classification import (accuracy, brier_score, compute_counts_confs_accs, ece,
                                                    expected_calibration_error, maximum_calibration_error, mce)

    r = np.random.default_rng(2)
    probs = P.softmax((3 * r.standard_normal((1234, 7))).astype(np.float32))
    targets = r.integers(0, 7, 1234)
    preds = probs.argmax(-1)
    counts, confs, accs = compute_counts_confs_accs(preds, probs, targets)
    oc, oconf, oacc, _ = Sc.compute_counts_confs_accs(preds, probs, targets)
    assert np.array_equal(counts.cpu().numpy(), oc)
    np.testing.assert_allclose(confs.cpu().numpy(), oconf, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(accs.cpu().numpy(), oacc, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(float(expected_calibration_error(preds, probs, targets)),
                               Sc.expected_calibration_error(preds, probs, targets), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(float(maximum_calibration_error(preds, probs, targets)),
                               Sc.maximum_calibration_error(preds, probs, targets), rtol=1e-5, atol=1e-6)
    assert float(ece(preds, probs, targets)) == float(expected_calibration_error(preds, probs, targets))
    assert float(mce(preds, probs, targets)) == float(maximum_calibration_error(preds, probs, targets))
    np.testing.assert_allclose(float(accuracy(preds, targets)), Sc.accuracy(preds, targets), rtol=1e-6)
    np.testing.assert_allclose(float(brier_score(probs, targets)), Sc.brier_score(probs, targets), rtol=1e-5)
    assert np.atleast_1d(brier_score(probs, targets).cpu().numpy()).shape == (1,)  # reference test_metric.py
    with pytest.raises(ValueError):
        accuracy(np.zeros((3, 2), np.int32), targets[:3])
    with pytest.raises(ValueError):
        brier_score(np.zeros((2, 2, 2), np.float32), targets[:2])
    # precomputed (one-dimensional) confidences, as the reference allows
    conf = probs.max(-1)
    c2, _, _ = compute_counts_confs_accs(preds, conf, targets)
    assert np.array_equal(c2.cpu().numpy(), oc)
