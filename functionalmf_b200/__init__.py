"""B200-native Gibbs-sweep engine for Bayesian Tensor Filtering.

Drop-in for the sampler loop of tansey/functionalmf: the same model classes and
``run_gibbs()`` result dict, every numerical step in hand-written sm_100a CUDA
behind the C ABI of ``include/btf_b200.h`` (``libbtf_b200.so``).
"""
from .factor import (BayesianTensorFiltering, GaussianBayesianTensorFiltering,          # noqa: F401
                     BinomialBayesianTensorFiltering, NegativeBinomialBayesianTensorFiltering)
from .constrained import (ConstrainedNonconjugateBayesianTensorFiltering,                # noqa: F401
                          NonconjugateBayesianTensorFiltering)
from ._lib import BTFError, BTFLibraryError, NotPositiveDefiniteError                    # noqa: F401
from .metrics import HeldOutEvaluator, heldout_classes                                      # noqa: F401
from .datasets import load_politics, load_flu_states                                     # noqa: F401
from .utils import ilogit, mse, mae, bayes_grid_penalty                                  # noqa: F401

__version__ = '0.1.0'
