"""Parallel-tempering (Langevin) Monte Carlo driver: the sampler behind the reference's
`Chain.samplerPTLMC` / `tempexchange` / `run_MCMC_PTLMC` (src/mcmc.py:431-727; the reference took it
from surmise 0.2.1's PTLMC and adapted the log-posterior signature and the chain layout).

It is a caller of the hot path, not part of it: every iteration evaluates all `numtemps + numchain`
chains with ONE `logpostfunc(theta[n, p])` call (the GPU log-posterior at n ~ 66), and the
pre-optimiser makes N = 1 calls from L-BFGS-B.  The implementation is organised differently from the
reference (ladder / pre-optimiser / move / exchange are separate functions) but performs the same
operations in the same order and draws from NumPy's global generator in the same sequence, so a
seeded run reproduces the reference's chain (tests/test_ptlmc.py, golden from the unmodified
reference).

Algorithm (no-gradient branch is what Chain.log_posterior exercises):
  ladder      temps = [exp(linspace(log T, log T / (numtemps + 1), numtemps)), 1 x numchain]
  start       rank the candidate points by -lp + p * N(0,1)^2, keep the best numtemps + numchain,
              polish each with L-BFGS-B in standardised coordinates, then kick it off the optimum
              along the inverse-Hessian metric (step 4, halved down to 1/16 while lp drops > 3p)
  move        theta' = theta + sqrt(2) rho_T (xi @ C^1/2)   [+ rho_T^2 grad @ C with gradients],
              rho_T = rho T^(1/3);  Metropolis on lp / T (with the Langevin correction if gradients)
  exchange    5 sweeps of random adjacent-temperature swaps (PT rule)
  tuning      every 10 steps of the first 2 x sampperchain: tau += (acc/10 - target) / sqrt(1 + k/10),
              rho = 2 (1 + tanh tau); target 0.25 (0.60 with gradients)
  output      theta[numchain, sampperchain, p]: the T = 1 chains after tuning
"""
from __future__ import annotations

import logging

import numpy as np
import scipy.optimize as spo

log = logging.getLogger(__name__)


def temperature_ladder(numtemps, numchain, maxtemp):
    """column vector [numtemps + numchain, 1]: geometric ladder from maxtemp down, then ones"""
    hot = np.exp(np.linspace(np.log(maxtemp), np.log(maxtemp) / (numtemps + 1), numtemps))
    return np.array(np.concatenate((hot, np.ones(numchain))), ndmin=2).T


def temp_exchange(lpostf, temps, iters=1):
    """Random adjacent swaps along the ladder; returns the new ordering of the chains.
    lpostf [n] or [n, 1] untempered log-posteriors, temps likewise (src/mcmc.py:679-693).

    Per sweep the reference draws n slots, then one uniform per slot as it walks through them; n
    uniforms drawn in one call are the same numbers in the same order, so the draws are made up
    front and the (inherently sequential) swap loop runs in the library's host helper."""
    from . import _lib
    n = lpostf.shape[0]
    lp = np.ascontiguousarray(lpostf, dtype=np.float64).reshape(n)
    tt = np.ascontiguousarray(temps, dtype=np.float64).reshape(n)
    order = np.arange(0, n, dtype=np.int64)
    slots = np.arange(1, n)
    for _ in range(iters):
        picks = np.ascontiguousarray(np.random.choice(slots, n), dtype=np.int64)
        log_u = np.log(np.random.uniform(size=n))
        _lib.check(_lib.lib.gpbt_host_temp_exchange(_lib.host_ptr(lp), _lib.host_ptr(tt), n, _lib.host_ptr(picks),
                                                    _lib.host_ptr(log_u), n, _lib.host_ptr(order)))
    return order


def temp_exchange_python(lpostf, temps, iters=1):
    """The same sweep as a plain Python loop (kept for the tests: draw-for-draw the reference's)."""
    n = lpostf.shape[0]
    order = np.arange(0, n)
    slots = np.arange(1, n)
    for _ in range(iters):
        for rt in np.random.choice(slots, n):
            gap = 1 / temps[rt - 1] - 1 / temps[rt]
            if (lpostf[order[rt]] - lpostf[order[rt - 1]]) * gap > np.log(np.random.uniform(size=1)):
                order[rt - 1], order[rt] = order[rt], order[rt - 1]
    return order


class _Target:
    """Normalises the three call conventions the reference accepts: plain lp, (lp, grad) tuples, and
    tuples that can be switched off with return_grad=False."""

    def __init__(self, fn, probe2, probe1):
        self.fn = fn
        out = fn(probe2)
        self.has_grad = type(out) is tuple
        self._kw = False
        if self.has_grad:
            if len(out) != 2:
                raise ValueError("log density does not return 1 or 2 elements")
            if out[1].shape[1] != probe2.shape[1]:
                raise ValueError("derivative appears to be the wrong shape")
            try:
                if type(fn(probe1, return_grad=False)) is tuple:
                    raise ValueError("Cannot stop returning a grad")
                self._kw = True
            except Exception:
                self._kw = False

    def value(self, theta):
        """lp as a column [n, 1]"""
        if not self.has_grad:
            return np.array(self.fn(theta), ndmin=2).T
        return self.fn(theta, return_grad=False) if self._kw else self.fn(theta)[0]

    def both(self, theta):
        return self.fn(theta)

    def grad(self, theta):
        return self.fn(theta)[1]


def _preoptimise(target, theta, cen, sc):
    """L-BFGS-B polish of every starting point in standardised coordinates, then a kick along the
    inverse-Hessian metric (rows of theta are updated in place)."""
    p = theta.shape[1]

    def neg(tp):
        x = cen + sc * tp
        return -target.value(x.reshape((1, len(x))))[0]

    jac = None
    if target.has_grad:
        def jac(tp):
            x = cen + sc * tp
            return -sc * target.grad(x.reshape((1, len(x))))

    z = (theta - cen) / sc
    box = spo.Bounds(np.maximum(-10 * np.ones(p), np.min(z, 0)), np.minimum(10 * np.ones(p), np.max(z, 0)))
    for k in range(theta.shape[0]):
        if k % 10 == 0:
            log.info("PTLMC pre-optimisation: chain %d", k)
        res = spo.minimize(neg, (theta[k, :] - cen) / sc, method="L-BFGS-B", jac=jac, bounds=box)
        theta[k, :] = cen + sc * res.x
        if k == 0:
            continue                     # the first chain stays on its optimum
        W, V = np.linalg.eigh(res.hess_inv @ np.eye(p))
        base, stride = neg(res.x), 4
        while True:
            kick = (V.T * np.sqrt(W)) @ (V @ np.random.standard_normal(size=p))
            if (neg(stride * kick + res.x) - base) < 3 * p:
                theta[k, :] = cen + sc * (stride * kick + res.x)
                break
            stride /= 2
            if stride < 1 / 16:
                theta[k, :] = cen + sc * res.x
                break
    return theta


class DevicePTLMC:
    """The iteration loop of the sampler with the chains resident on the GPU (gpbt_ptlmc_*, csrc/ptlmc.cuh): Gaussian
    proposal, the chain's log-posterior path, tempered Metropolis accept, five exchange sweeps, step-size tuning and
    the record of the T = 1 chains, without a host round trip per iteration.  Same algorithm as the loop of
    `sampler_ptlmc` (src/mcmc.py:623-671, 679-693); the random numbers come from a counter-based Philox generator
    keyed by `seed`, so a run is reproducible but does not follow NumPy's global stream (the host loop does).

    temps: the ladder [n] (hot chains first, the T = 1 chains last), root [p, p] = C^1/2 of the proposal,
    n_hot = number of chains above T = 1."""

    def __init__(self, device_chain, temps, root, n_hot, goal=0.25, seed=None):
        import ctypes as C
        from . import _lib
        self._lib = _lib
        temps = np.ascontiguousarray(temps, dtype=np.float64).reshape(-1)
        root = np.ascontiguousarray(root, dtype=np.float64)
        self.n, self.p, self.n_hot = temps.shape[0], device_chain.p, int(n_hot)
        if root.shape != (self.p, self.p):
            raise ValueError("root must be [%d, %d]" % (self.p, self.p))
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        self.seed = int(seed)
        if not device_chain._checked:
            device_chain._self_check()
        h = C.c_void_p()
        _lib.check(_lib.lib.gpbt_ptlmc_create(C.byref(h), device_chain.handle(), self.n, self.n_hot, _lib.host_ptr(temps),
                                              _lib.host_ptr(root), float(goal), C.c_uint64(self.seed)))
        self._h, self._dc = h, device_chain
        device_chain._dependents.add(self)
        self.n_tune = self.n_keep = 0

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._lib.lib.gpbt_ptlmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass   # interpreter shutdown

    def _handle(self):
        if self._h is None:
            raise RuntimeError("the sampler was closed (its chain was released or rebuilt)")
        return self._h

    def set_state(self, theta, tau=-1.0):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        if theta.shape != (self.n, self.p):
            raise ValueError("theta must be [%d, %d]" % (self.n, self.p))
        if not np.all(np.isfinite(theta)):
            raise ValueError("the starting points contain non-finite coordinates")
        self._lib.check(self._lib.lib.gpbt_ptlmc_set_state(self._handle(), self._lib.host_ptr(theta), float(tau)))

    def run(self, n_tune, n_keep, n_steps=None):
        """the next n_steps (default: all) iterations of a run of n_tune tuning + n_keep recorded iterations"""
        self.n_tune, self.n_keep = int(n_tune), int(n_keep)
        if n_steps is None:
            n_steps = self.n_tune + self.n_keep
        self._lib.check(self._lib.lib.gpbt_ptlmc_run(self._handle(), self.n_tune, self.n_keep, int(n_steps)))

    def read(self):
        """{'theta': [n - n_hot, n_keep, p] recorded T = 1 chains, 'state' [n, p], 'lp' [n], 'tau', 'stride',
        'accepted' (proposals of the T = 1 chains taken after tuning), 'n_notpd'}"""
        saved = np.zeros((self.n - self.n_hot, self.n_keep, self.p))
        state, lp, info = np.empty((self.n, self.p)), np.empty(self.n), np.zeros(4)
        hp = self._lib.host_ptr
        self._lib.check(self._lib.lib.gpbt_ptlmc_read(self._handle(), hp(saved) if saved.size else None, hp(state), hp(lp),
                                                      hp(info)))
        return {"theta": saved, "state": state, "lp": lp, "tau": info[0], "stride": info[1], "accepted": int(info[2]),
                "n_notpd": int(info[3])}


def sampler_ptlmc(logpostfunc, draw_func, theta0=None, numtemps=32, numchain=16, sampperchain=400,
                  maxtemp=30, nstartparameters=1000, exchange=temp_exchange, device_chain=None, seed=None):
    """Returns {'theta': [numchain, sampperchain, p]} (src/mcmc.py:431-675).

    device_chain (a DeviceChain whose log-posterior logpostfunc evaluates): run the iteration loop on the GPU
    (DevicePTLMC, Philox draws keyed by `seed`) after the start-up stage, which stays on the host either way."""
    if theta0 is None:
        theta0 = draw_func(nstartparameters)
    if theta0.shape[0] < 10 * theta0.shape[1]:        # too few candidates to rank: draw (again)
        theta0 = draw_func(nstartparameters)
    p = theta0.shape[1]
    n_tune = np.ceil(sampperchain * 2.0).astype("int")
    n_all = numtemps + numchain
    temps = temperature_ladder(numtemps, numchain, maxtemp)
    target = _Target(logpostfunc, theta0[0:2, :], theta0[10, :])
    goal = 0.60 if target.has_grad else 0.25

    log.info("PTLMC: ranking %d starting points", theta0.shape[0])
    score = -np.squeeze(target.value(theta0)) + p * np.random.standard_normal(size=theta0.shape[0]) ** 2
    theta = theta0[np.argsort(score)[0:n_all], :]
    cen = np.mean(theta, 0)
    sc = np.maximum(np.std(theta, 0), 10 ** (-8) * np.std(theta))
    theta = _preoptimise(target, theta, cen, sc)

    if target.has_grad:
        f, df = target.both(theta)
        f, df = f / temps, df / temps
    else:
        f, df = target.value(theta) / temps, None
    saved = np.zeros((numchain, sampperchain, p))
    cov = np.cov(theta.T)
    if p > 1:
        cov = 0.9 * cov + 0.1 * np.diag(np.diag(cov))       # keeps every direction moving
        W, V = np.linalg.eigh(cov)
        root = V @ np.diag(np.sqrt(W)) @ V.T
    else:
        root = np.sqrt(cov).reshape(1, 1)
        cov = cov.reshape(1, 1)

    def stride_of(tau):
        return 2 * (1 + (np.exp(2 * tau) - 1) / (np.exp(2 * tau) + 1))

    tau = -1
    if device_chain is not None:
        if target.has_grad:
            raise ValueError("the device loop is the sampler's branch without gradients")
        dev = DevicePTLMC(device_chain, temps, root, numtemps, goal=goal, seed=seed)
        try:
            dev.set_state(theta, tau=tau)
            dev.run(n_tune, sampperchain)
            out = dev.read()
        finally:
            dev.close()
        log.info("PTLMC (device loop): tau %.3f, %d accepted proposals of the T = 1 chains in %d iterations",
                 out["tau"], out["accepted"], sampperchain)
        return {"theta": out["theta"]}
    rho_t = stride_of(tau) * temps ** (1 / 3)
    hits = 0
    for k in range(0, n_tune + sampperchain):
        if k % 100 == 0:
            log.info("PTLMC: iteration %d", k)
        xi = np.random.normal(0, 1, theta.shape)
        prop = theta + np.sqrt(2) * rho_t * (xi @ root)
        if target.has_grad:
            prop += (rho_t ** 2) * (df @ cov)
            fp, dfp = target.both(prop)
            fp, dfp = fp / temps, dfp / temps
            a = xi / np.sqrt(2)
            b = (rho_t / 2) * ((df + dfp) @ root)
            corr = -(2 * np.sum(a * b, 1) + np.sum(b ** 2, 1))
        else:
            fp = target.value(prop) / temps
            corr = np.zeros(fp.shape)
        draw = np.log(np.random.uniform(size=f.shape[0]))
        take = np.where(np.squeeze(draw) < np.squeeze(fp - f) + np.squeeze(corr))[0]
        if take.shape[0] > 0:
            hits = hits + take.shape[0] / n_all
            theta[take, :] = 1 * prop[take, :]
            f[take] = 1 * fp[take]
            if target.has_grad:
                df[take, :] = 1 * dfp[take, :]
        flat = f * temps
        order = exchange(flat, temps, iters=5)
        f = flat[order] / temps
        theta = theta[order, :]
        if target.has_grad:
            df = (1 / temps) * (temps * df)[order, :]
        if k < n_tune and k % 10 == 0:
            tau = tau + 1 / np.sqrt(1 + k / 10) * ((hits / 10) - goal)
            rho_t = stride_of(tau) * (temps ** (1 / 3))
            hits = 0
        elif k >= n_tune:
            saved[:, k - n_tune, :] = 1 * theta[numtemps:, ]
    return {"theta": saved}
