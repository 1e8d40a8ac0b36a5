"""A minimal stand-in for emcee (not installed here): an affine-invariant stretch-move ensemble
sampler exposing the part of emcee 3's EnsembleSampler API that Chain.run_mcmc uses.  Like emcee,
it evaluates the log-probability of a whole half-ensemble through `pool.map(log_prob_fn, coords)`,
which is the call the reference hijacks with pool=self (src/mcmc.py:335-342, 372-374)."""
import numpy as np


class State:
    def __init__(self, coords, log_prob):
        self.coords, self.log_prob = coords, log_prob

    def __iter__(self):   # emcee's State unpacks to (coords, log_prob, random_state)
        return iter((self.coords, self.log_prob, None))


class EnsembleSampler:
    def __init__(self, nwalkers, ndim, log_prob_fn, pool=None, a=2.0, seed=0):
        self.nwalkers, self.ndim, self.log_prob_fn, self.pool, self.a = nwalkers, ndim, log_prob_fn, pool, a
        self.rng = np.random.default_rng(seed)
        self.n_calls = 0
        self.reset()

    def reset(self):
        self._chain, self._lp = [], []
        self._accepted = np.zeros(self.nwalkers)
        self._iters = 0

    def _logp(self, p):
        self.n_calls += 1
        mapper = self.pool.map if self.pool is not None else map
        out = np.array(list(mapper(self.log_prob_fn, p)), dtype=np.float64)
        if np.any(np.isnan(out)):
            raise ValueError("Probability function returned NaN")
        return out

    @property
    def acceptance_fraction(self):
        return self._accepted / max(self._iters, 1)

    def get_chain(self, flat=False):
        c = np.array(self._chain)                       # [step, walker, dim]
        return c.reshape(-1, self.ndim) if flat else c

    def get_log_prob(self, flat=False):
        lp = np.array(self._lp)
        return lp.reshape(-1) if flat else lp

    def sample(self, initial_state, iterations=1, skip_initial_state_check=False):
        x = np.array(initial_state.coords if isinstance(initial_state, State) else initial_state, dtype=np.float64)
        lp = self._logp(x)
        half = self.nwalkers // 2
        for _ in range(iterations):
            for first in (True, False):
                s = slice(0, half) if first else slice(half, self.nwalkers)
                c = slice(half, self.nwalkers) if first else slice(0, half)
                ns = x[s].shape[0]
                z = ((self.a - 1.0) * self.rng.random(ns) + 1.0) ** 2 / self.a
                partner = x[c][self.rng.integers(0, x[c].shape[0], ns)]
                prop = partner + z[:, None] * (x[s] - partner)
                lp_new = self._logp(prop)
                log_ratio = (self.ndim - 1) * np.log(z) + lp_new - lp[s]
                acc = np.log(self.rng.random(ns)) < log_ratio
                xs, ls = x[s].copy(), lp[s].copy()
                xs[acc], ls[acc] = prop[acc], lp_new[acc]
                x[s], lp[s] = xs, ls
                self._accepted[s] += acc
            self._iters += 1
            self._chain.append(x.copy())
            self._lp.append(lp.copy())
            yield State(x.copy(), lp.copy())
