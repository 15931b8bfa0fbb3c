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
from .multivalid.iterative.classification.binary_multicalibrator import BinaryClassificationMulticalibrator  # noqa: F401
from .multivalid.iterative.classification.top_label_multicalibrator import TopLabelMulticalibrator  # noqa: F401
from .multivalid.iterative.multicalibrator import Multicalibrator  # noqa: F401
from .multivalid.iterative.regression.batch_mvp import BatchMVPConformalClassifier  # noqa: F401
from .multivalid.one_shot.classification.binary_multicalibrator import (  # noqa: F401
    OneShotBinaryClassificationMulticalibrator,
)
from .multivalid.one_shot.classification.top_label_multicalibrator import OneShotTopLabelMulticalibrator  # noqa: F401
from .multivalid.one_shot.multicalibrator import OneShotMulticalibrator  # noqa: F401
from .regression.batch_mvp import BatchMVPConformalRegressor  # noqa: F401
