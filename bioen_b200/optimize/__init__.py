"""Drop-in mirror of the reference package `bioen.optimize` (bioen/optimize/__init__.py), GPU backed."""
from . import common  # noqa: F401
from . import forces  # noqa: F401
from . import log_weights  # noqa: F401
from . import minimize  # noqa: F401
from . import util  # noqa: F401
