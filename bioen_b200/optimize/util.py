"""General utilities of the optimize API (mirror of bioen/optimize/util.py)."""
import numpy as np
import yaml

from .ext import c_bioen


def library_gsl():
    """True: the GSL-style minimisers are built into libbioen_b200.so (bioen/optimize/util.py:13-25)."""
    return c_bioen.library_gsl()


def library_lbfgs():
    """True: the liblbfgs-style minimiser is built into libbioen_b200.so (bioen/optimize/util.py:28-40)."""
    return c_bioen.library_lbfgs()


def compute_relative_difference_for_values(a, b):
    """|a-b|/|b|, or |a| when b == 0 (bioen/optimize/util.py:43-64)."""
    if b == 0:
        return abs(a)
    return abs(a - b) / abs(b)


def compute_relative_difference_for_arrays(a, b):
    """Largest element-wise relative difference over the non-zero entries of b and its position
    (bioen/optimize/util.py:67-92)."""
    a = np.asarray(a)
    b = np.asarray(b)
    msk = b != 0.0
    if not msk.any():
        return 0.0, 0
    d_array = np.abs(a[msk] - b[msk]) / np.abs(b[msk])
    idx = int(np.argmax(d_array))
    return d_array[idx], idx


def ntype(s):
    """String -> int | float | bool | str (bioen/optimize/util.py:163-188)."""
    for conv in (int, float):
        try:
            return conv(s)
        except Exception:
            pass
    low = s.lower()
    if low in ("true", "t", "yes", "y"):
        return True
    if low in ("false", "f", "no", "n"):
        return False
    return s


def nested_set(dic, keys, value):
    """dic[k0][k1]...[kn] = value, creating levels as needed (bioen/optimize/util.py:191-196)."""
    for key in keys[:-1]:
        dic = dic.setdefault(key, {})
    dic[keys[-1]] = value


def load_template_config_yaml(file_name, minimizer, parameter_mod=""):
    """Build the flat cfg dict for `minimizer` from the yaml template, applying "sect:key=value,..." overrides
    (bioen/optimize/util.py:95-160).  Keys: minimizer, debug, verbose, params, n_threads,
    cache_ytilde_transposed, algorithm, use_c_functions."""
    minimizer = minimizer.lower()
    with open(file_name, "r") as fp:
        cfg = yaml.safe_load(fp)
    if parameter_mod:
        for token in parameter_mod.split(","):
            keys, value = token.split("=")
            nested_set(cfg, keys.split(":"), ntype(value))
    params = dict(cfg[minimizer])
    packed = {
        "minimizer": minimizer,
        "debug": cfg["general"]["debug"],
        "verbose": cfg["general"]["verbose"],
        "params": params,
        "n_threads": cfg["c_functions"]["n_threads"],
        "cache_ytilde_transposed": cfg["c_functions"]["cache_ytilde_transposed"],
    }
    packed["algorithm"] = params.pop("algorithm", "")
    packed["use_c_functions"] = params.pop("use_c_functions", True)
    return packed
