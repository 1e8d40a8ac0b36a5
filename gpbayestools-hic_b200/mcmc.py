"""
Drop-in for the reference's `src/mcmc.py` posterior layer: `Chain.log_posterior`,
`Chain.log_likelihood`, `Chain._predict`, `mvn_loglike` keep their signatures and return
conventions, so emcee (pool=self -> Chain.map), pocoMC (likelihood=chain.log_likelihood,
likelihood_kwargs={'finite': True}, vectorize=True) and PTLMC (logpostfunc=chain.log_posterior)
call them unchanged -- but every evaluation runs on the B200 through the C ABI in include/gpbt.h.

Reference lines this mirrors: constructor src/mcmc.py:104-142, loadEmulator :145-150,
_predict :153-166, log_prior :169-185, log_likelihood :188-222, log_likelihood_point_by_point
:225-258, log_posterior :261-299, _read_in_exp_data_pickle :302-324, random_pos :327-332,
map :335-342, run_mcmc :345-426, compute_log_likelihood_for_chain :729-749, run_pocoMC :752-819.
"""
from __future__ import annotations

import logging
import pickle
from pathlib import Path

import numpy as np

from . import parse_model_parameter_file
from .device import DeviceChain, as_rows, default_devices, mvn_loglike_batch
from .state import EmulatorState

log = logging.getLogger(__name__)


def mvn_loglike(y, cov):
    """Unnormalised multivariate-normal log density  -1/2 y^T C^-1 y - 1/2 log det C  for one
    difference vector `y` [n] and covariance `cov` [n, n] (src/mcmc.py:23-65), evaluated by the
    batched Cholesky kernel.  A non-positive-definite `cov` raises LinAlgError (the reference
    intends to, src/mcmc.py:49-54, but tests the wrong sign of `info`)."""
    y = np.asarray(y, dtype=np.float64).reshape(1, -1)
    val = mvn_loglike_batch(y, np.asarray(cov, dtype=np.float64).reshape(1, y.shape[1], y.shape[1]),
                            notpd_value=np.nan)[0]
    if np.isnan(val):
        raise np.linalg.LinAlgError("covariance is not positive definite")
    return float(val)


def _dist_world():
    """(world size, rank) of the default torch.distributed group, (1, 0) when there is none"""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
    except ImportError:
        pass
    return 1, 0


def read_experiment_pickle(path):
    """Experimental data in the training-pickle layout (src/mcmc.py:302-324): returns the values
    [n_entries, m] and the diagonal covariance diag(err^2) [m, m] of the flattened errors."""
    with open(path, "rb") as fh:
        entries = pickle.load(fh)
    vals = np.array([np.asarray(e["obs"], dtype=np.float64)[0] for e in entries.values()])
    errs = np.nan_to_num(np.abs(np.array([np.asarray(e["obs"], dtype=np.float64)[1]
                                          for e in entries.values()]))).ravel()
    return vals, np.diag(errs ** 2)


class Chain:
    """Posterior evaluation + sampler drivers with the reference's interface."""

    def __init__(self, mcmc_path="./mcmc/chain.pkl", expdata_path="./exp_data.dat",
                 model_parafile="./model.dat"):
        self.mcmc_path = Path(mcmc_path)
        self.mcmc_path.parent.mkdir(exist_ok=True)
        self.pardict = parse_model_parameter_file(model_parafile)
        self.ndim = len(self.pardict)
        self.label = [v[0] for v in self.pardict.values()]
        self.min = np.array([v[1] for v in self.pardict.values()], dtype=np.float64)
        self.max = np.array([v[2] for v in self.pardict.values()], dtype=np.float64)
        self.prior_volume_ = np.prod(self.max - self.min)
        self.expdata, self.expdata_cov = read_experiment_pickle(expdata_path)
        self.nobs = self.expdata.shape[1]
        self.emuList = []
        self.chain = False
        # GPUs large host batches are spread over: None = automatic (gpbt_b200.device.default_devices: all
        # visible GPUs of a single-process run, the rank's own GPU under torchrun), or a list of indices
        self.devices = None
        self._dev = None
        self._dev_key = None
        self._dev_snapshot = None
        self._dev_calls = 0

    # ---- emulators -----------------------------------------------------------------------
    def loadEmulator(self, emulatorPathList):
        """dill-load trained emulators (this package's Emulator or a reference
        src.emulator.Emulator pickle -- anything EmulatorState.from_trained understands)."""
        import dill
        for path in emulatorPathList:
            with open(path, "rb") as fh:
                self.emuList.append(dill.load(fh))
        self._dev = None
        log.info("Number of Emulators: %d", len(self.emuList))

    def _states(self):
        out = []
        for emu in self.emuList:
            if isinstance(emu, EmulatorState):
                out.append(emu)
            elif isinstance(getattr(emu, "state", None), EmulatorState):
                out.append(emu.state)
            else:
                st = getattr(emu, "_gpbt_state", None)
                if st is None:
                    st = EmulatorState.from_trained(emu)
                    try:
                        emu._gpbt_state = st
                    except AttributeError:
                        pass
                out.append(st)
        return out

    _COV_SAMPLE = 97          # stride of the sampled elements of expdata_cov in the per-call check
    _FULL_CHECK_EVERY = 256   # calls between full comparisons of expdata_cov

    def invalidate(self):
        """Drop the device-resident copy of the chain; the next call re-uploads emulators, bounds and
        experimental data.  Needed only after edits the automatic check cannot see (see device())."""
        if self._dev is not None:
            self._dev.release()
        self._dev = None
        self._dev_key = None

    def _state_key(self, full):
        """What the device copy was built from.  The reference re-reads min / max / expdata /
        expdata_cov on every call (src/mcmc.py:188-222), so in-place edits must not leave a stale copy
        on the GPU: the small arrays are compared in full on every call, expdata_cov (nobs^2 values) by
        its diagonal and a strided sample on every call and in full every _FULL_CHECK_EVERY calls (and
        whenever `full` is set: large batches, where the comparison costs nothing)."""
        cov = np.asarray(self.expdata_cov)
        key = [tuple(id(e) for e in self.emuList), cov.shape,
               np.asarray(self.min, dtype=np.float64).tobytes(), np.asarray(self.max, dtype=np.float64).tobytes(),
               np.asarray(self.expdata, dtype=np.float64).tobytes(),
               np.ascontiguousarray(cov.diagonal()).tobytes(), cov.ravel()[::self._COV_SAMPLE].tobytes(),
               None if self.devices is None else tuple(self.devices)]
        snap = self._dev_snapshot
        if full and snap is not None and cov.shape == snap.shape and not np.array_equal(cov, snap):
            key.append("expdata_cov changed")
        return tuple(key)

    def device(self, rows=0) -> DeviceChain:
        """Device-resident chain; rebuilt when emuList, the bounds, expdata or expdata_cov change --
        by assignment or in place (content check, see _state_key; `invalidate()` forces it)."""
        self._dev_calls += 1
        full = rows >= 1024 or self._dev_calls % self._FULL_CHECK_EVERY == 0
        key = self._state_key(full)
        if self._dev is None or key != self._dev_key:
            if not self.emuList:
                raise RuntimeError("no emulator loaded (call loadEmulator first)")
            if self._dev is not None:
                self._dev.release()
            devices = self.devices if self.devices is not None else default_devices()
            self._dev = DeviceChain(self._states(), self.min, self.max, self.expdata[0], self.expdata_cov,
                                    devices=devices)
            if self._dev.M != self.nobs:
                raise ValueError("emulators predict %d observables, experiment has %d" % (self._dev.M, self.nobs))
            self._dev_snapshot = np.array(self.expdata_cov, dtype=np.float64, copy=True)
            self._dev_key = self._state_key(False)
        return self._dev

    # ---- hot path ------------------------------------------------------------------------
    def _predict(self, X, extra_std=0.0):
        """Concatenated emulator means [N, nobs] and block-diagonal covariance [N, nobs, nobs];
        every emulator sees extra_std * X[:, -1] as its extra_std array."""
        return self.device().predict(as_rows(X, self.ndim), extra_std=float(extra_std))

    def log_prior(self, X):
        """Normalised uniform prior on the box; -inf outside (not used by log_posterior, as in
        the reference)."""
        X = as_rows(X, self.ndim)
        lp = np.full(X.shape[0], -np.log(self.prior_volume_))
        lp[~np.all((X > self.min) & (X < self.max), axis=1)] = -np.inf
        return lp

    def log_likelihood(self, X, extra_std_prior_scale=0.001, finite=False):
        """log L at each row of X.  Rows outside the (open) parameter box get -inf, or -1e300 with
        `finite=True` (what pocoMC needs).  `extra_std_prior_scale` is accepted for signature
        compatibility: the reference multiplies its extra_std by 0.0, which leaves the constant
        2*log(1e-16) and nothing that depends on the scale (src/mcmc.py:199-221)."""
        X = as_rows(X, self.ndim)
        return self.device(X.shape[0]).log_target(X, -1e300 if finite else -np.inf)

    def log_posterior(self, X, extra_std_prior_scale=.05):
        """Posterior at each row of X; identical to log_likelihood(finite=False): the reference
        adds no log-prior inside the box (src/mcmc.py:261-299).  Large batches are spread over the GPUs
        of `self.devices` from this one process (gpbt_fanout_log_posterior_host)."""
        X = as_rows(X, self.ndim)
        return self.device(X.shape[0]).log_target(X, -np.inf)

    def log_likelihood_point_by_point(self, X, extra_std_prior_scale=0.001):
        """Same values as the reference's N=1 Python loop (src/mcmc.py:225-258), in one batch."""
        X = as_rows(X, self.ndim)
        return self.device(X.shape[0]).log_target(X, -np.inf)

    def _read_in_exp_data_pickle(self, filepath):
        """values [n_entries, nobs] and diag(err^2) [nobs, nobs] of an experiment pickle (src/mcmc.py:302-324)"""
        return read_experiment_pickle(filepath)

    def random_pos(self, n=1):
        return np.random.uniform(self.min, self.max, (n, self.ndim))

    @staticmethod
    def map(f, args):
        """Lets a Chain stand in as emcee's `pool`: the whole walker array goes to `f` at once."""
        return f(args)

    def compute_log_likelihood_for_chain(self, output_path="./mcmc/log_likelihood.pkl"):
        if self.chain is False:
            with open(self.mcmc_path, "rb") as fh:
                self.chain = pickle.load(fh)["chain"]
        flat = np.ascontiguousarray(self.chain.reshape(-1, self.ndim))
        ll = self.log_likelihood_point_by_point(flat).reshape(self.chain.shape[0], self.chain.shape[1])
        with open(output_path, "wb") as fh:
            pickle.dump({"log_likelihood": ll}, fh)
        return ll

    # ---- sampler drivers (callers of the hot path; third-party samplers imported lazily) ----
    def run_mcmc(self, nsteps=500, nburnsteps=None, nwalkers=None, status=None, nthin=10,
                 skip_initial_state_check=False, sampler="emcee", seed=None):
        """Affine-invariant ensemble run with the reference's burn-in recipe: half the burn-in from
        random positions, restart from the best distinct points, second half, then production; the
        thinned chain is appended to `mcmc_path` (src/mcmc.py:345-426).

        sampler="emcee" (default, what the reference does, src/mcmc.py:372-374) drives
        emcee.EnsembleSampler with pool=self: one host call of log_posterior per half-ensemble;
        sampler="device" (opt-in) keeps the walkers on the GPU for the whole run
        (gpbt_b200.sampler.DeviceEnsembleSampler: the same stretch move, one CUDA graph per step,
        Philox random numbers keyed by `seed` -- statistically, not draw-for-draw, emcee)."""
        if nburnsteps is None or nwalkers is None:
            log.error("must specify nburnsteps and nwalkers to start chain")
            return
        stored = {}
        if self.mcmc_path.exists():
            with open(self.mcmc_path, "rb") as fh:
                stored = pickle.load(fh)
        world, rank = _dist_world()
        start = None
        if sampler == "device" and world > 1:
            # one process per GPU (torchrun): replicated walkers, proposals evaluated in slices.  The
            # replicas must make identical proposals, so rank 0's seed and starting positions are
            # broadcast -- the ranks' NumPy generators need not agree.
            import torch.distributed as dist
            from .sampler import ShardedEnsembleSampler
            if seed is None:
                seed = int(np.random.randint(0, 2 ** 31 - 1))
            shared = [int(seed), self.random_pos(nwalkers) if "chain" not in stored else None]
            dist.broadcast_object_list(shared, src=0)
            seed, start = shared
            ens = ShardedEnsembleSampler(nwalkers, self.ndim, self.device(), seed=seed)
        elif sampler == "device":
            from .sampler import DeviceEnsembleSampler
            ens = DeviceEnsembleSampler(nwalkers, self.ndim, self.device(), seed=seed)
        if sampler == "device":

            def advance(x0, steps):
                return ens.run_mcmc(x0, steps, status=status, skip_initial_state_check=skip_initial_state_check)
        elif sampler == "emcee":
            import emcee
            ens = emcee.EnsembleSampler(nwalkers, self.ndim, self.log_posterior, pool=self)

            def advance(x0, steps):
                every = status or max(steps // 10, 1)
                state = None
                for i, state in enumerate(ens.sample(x0, iterations=steps,
                                                     skip_initial_state_check=skip_initial_state_check), start=1):
                    if i % every == 0 or i == steps:
                        af = ens.acceptance_fraction
                        log.info("step %d: acceptance fraction: mean %.4f, std %.4f, min %.4f, max %.4f",
                                 i, af.mean(), af.std(), af.min(), af.max())
                return state
        else:
            raise ValueError("sampler must be 'device' or 'emcee', got %r" % (sampler,))

        if "chain" in stored:
            x0 = stored["chain"][:, -1, :]
        else:
            first = nburnsteps // 2
            advance(self.random_pos(nwalkers) if start is None else start, first)
            best = np.unique(ens.get_log_prob(flat=True), return_index=True)[1][-nwalkers:]
            x0 = ens.get_chain(flat=True)[best]
            ens.reset()
            x0 = advance(x0, nburnsteps - first)
            ens.reset()
        advance(x0, nsteps)
        thinned = np.swapaxes(ens.get_chain(), 0, 1)[:, ::nthin, :]   # [walker, step, dim]
        self.acceptance_fraction_ = np.asarray(ens.acceptance_fraction)
        self.chain = np.concatenate((stored["chain"], thinned), axis=1) if "chain" in stored else thinned
        stored["chain"] = self.chain
        if rank == 0:   # under torchrun every rank holds the same chain; one of them writes it
            with open(self.mcmc_path, "wb") as fh:
                pickle.dump(stored, fh)
        if sampler == "device":
            ens.close()

    # ---- parallel tempering (src/mcmc.py:431-727) -----------------------------------------
    def samplerPTLMC(self, logpostfunc, draw_func, theta0=None, numtemps=32, numchain=16, sampperchain=400,
                     maxtemp=30, nstartparameters=1000, sampler="host", seed=None):
        """Parallel-tempering (Langevin) MCMC; returns {'theta': [numchain, sampperchain, p]}.
        See gpbt_b200.ptlmc for the algorithm.  sampler="host" (default) is the reference's loop, draw for draw
        from NumPy's global generator; sampler="device" keeps the chains on the GPU for the iteration loop
        (Philox draws keyed by `seed`; logpostfunc must be this chain's log_posterior)."""
        from .ptlmc import sampler_ptlmc
        if sampler not in ("host", "device"):
            raise ValueError("sampler must be 'host' or 'device'")
        dc = None
        if sampler == "device":
            dc = self.device(numtemps + numchain)
        return sampler_ptlmc(logpostfunc, draw_func, theta0=theta0, numtemps=numtemps, numchain=numchain,
                             sampperchain=sampperchain, maxtemp=maxtemp, nstartparameters=nstartparameters,
                             exchange=self.tempexchange, device_chain=dc, seed=seed)

    def tempexchange(self, lpostf, temps, iters=1):
        from .ptlmc import temp_exchange
        return temp_exchange(lpostf, temps, iters=iters)

    def run_MCMC_PTLMC(self, nsteps=500, nwalkers=16, ntemps=50, maxtemp=100, nstartparameters=1000,
                       sampler="host", seed=None):
        """PTLMC run on the GPU log-posterior (all ntemps + nwalkers chains in one call per iteration);
        the T = 1 chains are stored as chain[nwalkers, nsteps, ndim] in `mcmc_path`
        (src/mcmc.py:696-727).  sampler="device": see samplerPTLMC."""
        out = self.samplerPTLMC(logpostfunc=self.log_posterior, draw_func=self.random_pos, theta0=None,
                                numtemps=ntemps, numchain=nwalkers, sampperchain=nsteps, maxtemp=maxtemp,
                                nstartparameters=nstartparameters, sampler=sampler, seed=seed)
        self.chain = out["theta"].reshape((nwalkers, nsteps, self.ndim))
        with open(self.mcmc_path, "wb") as fh:
            pickle.dump({"chain": self.chain}, fh)

    def run_pocoMC(self, n_effective=1000, n_active=250, n_prior=2000, sample="tpcn", n_max_steps=200,
                   random_state=42, n_total=5000, n_evidence=5000, pool=None, prior=None):
        """pocoMC preconditioned Monte Carlo with a uniform prior on the box (src/mcmc.py:752-819).
        `pool` must stay None / 1: a CUDA context does not survive fork(), and one GPU already
        evaluates all active particles in a single call."""
        import pocomc
        from scipy.stats import uniform
        if pool not in (None, 0, 1):
            raise ValueError("pool > 1 would fork the CUDA context; the GPU path is already batched")
        if prior is None:
            prior = pocomc.Prior([uniform(lo, hi - lo) for lo, hi in zip(self.min, self.max)])
        elif prior.dim != self.ndim:
            raise ValueError("prior.dim does not match the model parameter space")
        sampler = pocomc.Sampler(prior=prior, likelihood=self.log_likelihood,
                                 likelihood_kwargs={"finite": True}, n_effective=n_effective,
                                 n_active=n_active, n_prior=n_prior, sample=sample,
                                 n_max_steps=n_max_steps, random_state=random_state, vectorize=True,
                                 pool=None)
        sampler.run(n_total=n_total, n_evidence=n_evidence)
        samples, weights, logl, logp = sampler.posterior()
        logz, logz_err = sampler.evidence()
        with open(self.mcmc_path, "wb") as fh:
            pickle.dump({"chain": samples, "weights": weights, "logl": logl, "logp": logp,
                         "logz": logz, "logz_err": logz_err}, fh)
