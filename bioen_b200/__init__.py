"""bioen_b200 -- B200-native (sm_100a CUDA) implementation of BioEn's optimisation hot path.

`bioen_b200.optimize` mirrors the public API of the reference package `bioen.optimize`
(log_weights.find_optimum / forces.find_optimum, cfg dicts, 'scipy' | 'gsl' | 'lbfgs' minimisers);
`bioen_b200.Problem` is the device-resident handle underneath it.
"""
__version__ = "0.1.0"

from .problem import Problem, device_count, pinned_empty  # noqa: F401
