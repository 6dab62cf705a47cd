"""Shared helpers of the parity tests: hyper-parameter bags, oracle parameter loading, comparisons."""
import numpy as np

from oracle.params import HP, init_params, ssrn_specs, text2mel_specs


def make_hp(**kw):
    """Oracle HP + product Hyperparams carry the same hot-path fields; tests use one object for both sides."""
    from ophelia_b200.configuration import default_hparams
    hp = default_hparams(**kw)
    return hp


def oracle_params(hp, model, seed=0, perturb=True):
    specs = text2mel_specs(hp) if model == "t2m" else ssrn_specs(hp)
    return init_params(specs, seed, perturb=perturb)


def maxabs(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
