"""
Golden vectors for the parallel-tempering sampler: runs the UNMODIFIED reference
(`Chain.samplerPTLMC` / `Chain.tempexchange`, /root/reference/src/mcmc.py:431-693) on a toy
log-posterior with a fixed NumPy seed.  Build-container only; writes tests/golden/ptlmc_toy.npz.

    python tests/golden/make_golden_ptlmc.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _import_reference  # noqa: E402

LO = np.array([-3.0, -2.0, -4.0])
HI = np.array([4.0, 5.0, 3.0])
MEAN = np.array([0.5, 1.0, -0.5])
COV = np.array([[1.0, 0.5, 0.0], [0.5, 1.5, -0.4], [0.0, -0.4, 0.8]])
PREC = np.linalg.inv(COV)


def toy_logpost(X):
    """correlated Gaussian inside a box, -inf outside (the shape of Chain.log_posterior)"""
    X = np.array(X, dtype=np.float64, ndmin=2)
    d = X - MEAN
    lp = -0.5 * np.einsum("ni,ij,nj->n", d, PREC, d)
    lp[~np.all((X > LO) & (X < HI), axis=1)] = -np.inf
    return lp


def toy_with_grad(X, return_grad=True):
    X = np.array(X, dtype=np.float64, ndmin=2)
    d = X - MEAN
    lp = (-0.5 * np.einsum("ni,ij,nj->n", d, PREC, d)).reshape(-1, 1)
    return (lp, -d @ PREC) if return_grad else lp


def draw(n):
    return np.random.uniform(LO, HI, (n, 3))


def main():
    _, Chain, _ = _import_reference()
    ch = Chain.__new__(Chain)          # samplerPTLMC only uses self.tempexchange
    out = {}
    np.random.seed(20261018)
    out["theta"] = ch.samplerPTLMC(toy_logpost, draw, theta0=None, numtemps=5, numchain=3, sampperchain=25,
                                   maxtemp=12, nstartparameters=80)["theta"]
    np.random.seed(7)
    out["theta_grad"] = ch.samplerPTLMC(toy_with_grad, draw, theta0=None, numtemps=4, numchain=2, sampperchain=15,
                                        maxtemp=8, nstartparameters=60)["theta"]
    np.random.seed(99)
    lp = np.random.normal(size=(9, 1)) * 3
    temps = np.array(np.concatenate((np.exp(np.linspace(np.log(20), np.log(20) / 7, 6)), np.ones(3))), ndmin=2).T
    out["ex_lp"], out["ex_temps"] = lp, temps
    np.random.seed(100)
    out["ex_order"] = ch.tempexchange(lp, temps, iters=4)
    np.savez_compressed(os.path.join(HERE, "ptlmc_toy.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
