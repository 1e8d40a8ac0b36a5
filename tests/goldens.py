"""Golden fixtures for the tests: thin view of gpbt_b200.fixtures (tests/golden/*.npz, produced by
make_golden.py from the unmodified reference)."""
import importlib.util
import os

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _package_module(name):
    """a module of the package loaded by path: usable without importing the package (whose __init__
    chain needs the built CUDA library)"""
    spec = importlib.util.spec_from_file_location("gpbt_" + name, os.path.join(_ROOT, "gpbayestools-hic_b200", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_fx = _package_module("fixtures")
GOLDEN_DIR = _fx.GOLDEN_DIR
SMALL_CASES = ["c1_rbf", "c1_matern", "c1_logexp", "c1_nopca", "c1_multi", "odd_shape", "p20_trafo"]
available, load, rebuild_L = _fx.available, _fx.load, _fx.rebuild_L
oracle_states = _fx.state_dicts


def synthetic_module():
    return _package_module("synthetic")


def cov_exp_sys(g):
    """expdata_cov + the PSD systematic term make_golden.py added for lp_posterior_sys."""
    return g["cov_exp"] + synthetic_module().systematic_cov(g["cov_exp"].shape[0])
