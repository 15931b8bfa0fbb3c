"""User-facing names of fortuna.prob_model for the SG-MCMC hot path."""
from .calib_config import Adam, CalibConfig, CalibOptimizer, CalibProcessor  # noqa: F401
from .calibration import ClassificationTemperatureScaler, RegressionTemperatureScaler  # noqa: F401
from .classification import ProbClassifier  # noqa: F401
from .fit_config import FitCheckpointer, FitConfig, FitHyperparameters, FitMonitor, FitOptimizer, FitProcessor  # noqa: F401
from .posterior.sgmcmc.cyclical_sgld.cyclical_sgld_approximator import CyclicalSGLDPosteriorApproximator  # noqa: F401
from .posterior.sgmcmc.sghmc.sghmc_approximator import SGHMCPosteriorApproximator  # noqa: F401
from .regression import ProbRegressor  # noqa: F401
from .prior import IsotropicGaussianPrior  # noqa: F401
