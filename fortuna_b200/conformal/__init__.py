"""The conformal / calibration scorers of the hot path, under the reference's names (fortuna/conformal/__init__.py)."""
from .classification.adaptive_prediction import (  # noqa: F401
    AdaptivePredictionConformalClassifier,
    CVPlusAdaptivePredictionConformalClassifier,
)
from .classification.maxcovfixprec_binary_classfication import (  # noqa: F401
    MaxCoverageFixedPrecisionBinaryClassificationCalibrator,
)
from .classification.simple_prediction import (  # noqa: F401
    CVPlusSimplePredictionConformalClassifier,
    SimplePredictionConformalClassifier,
)
from .regression.quantile import OneDimensionalUncertaintyConformalRegressor, QuantileConformalRegressor  # noqa: F401
