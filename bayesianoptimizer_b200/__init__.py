"""B200-native GP surrogate + acquisition hot path behind the BayesianOptimizer call surface."""
from ._lib import (ACQ_EI, ACQ_LOGEI, ACQ_MEAN, ACQ_UCB, ACQ_VAR, KERNEL_LINEAR_MATERN52, KERNEL_MATERN52, KERNEL_RBF,
                   BoLibraryError, BoSobol)
from .engine import BoError, GPEngine, NotPositiveDefiniteError
from .sobol import sobol_state

__all__ = ["GPEngine", "BoError", "BoLibraryError", "NotPositiveDefiniteError", "BoSobol", "sobol_state",
           "ACQ_EI", "ACQ_LOGEI", "ACQ_UCB", "ACQ_VAR", "ACQ_MEAN", "KERNEL_MATERN52", "KERNEL_RBF", "KERNEL_LINEAR_MATERN52"]
