"""Device-resident affine-invariant ensemble sampler behind the part of emcee's EnsembleSampler API
that the reference uses (LoggingEnsembleSampler.run_mcmc and Chain.run_mcmc, src/mcmc.py:68-92,
345-426: `run_mcmc`, `sample`, `reset`, `chain`, `flatchain`, `flatlnprobability`, `lnprobability`,
`acceptance_fraction`, `get_chain`, `get_log_prob`).

The walkers never leave the GPU: one sampler step is a CUDA graph (split, 2 x [stretch proposal,
log-posterior path, Metropolis accept], record) replayed by gpbt_ensemble_run; the history of every
step is kept in HBM and copied back when it is asked for.  The move is emcee's default
(StretchMove, a = 2, random red/blue split); random numbers come from a counter-based Philox
generator keyed by `seed`, so a run is reproducible but does not follow NumPy's global stream."""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np

from . import _lib
from .device import DeviceChain, as_rows

log = logging.getLogger(__name__)


class State:
    """What emcee's State carries for this use: coords [nwalkers, ndim] and log_prob [nwalkers];
    unpacks to (coords, log_prob, random_state) like emcee's."""

    def __init__(self, coords, log_prob):
        self.coords, self.log_prob, self.random_state = coords, log_prob, None

    def __iter__(self):
        return iter((self.coords, self.log_prob, self.random_state))


class DeviceEnsembleSampler:
    def __init__(self, nwalkers, ndim, device_chain: DeviceChain, a=2.0, seed=None, randomize_split=True,
                 use_graph=True):
        if ndim != device_chain.p:
            raise ValueError("ndim = %d, but the chain has %d parameters" % (ndim, device_chain.p))
        if nwalkers < 2:
            raise ValueError("need at least 2 walkers")
        if nwalkers < 2 * ndim:
            # emcee raises here unless live_dangerously is set; an ensemble of fewer than 2 * ndim walkers
            # samples only the subspace it spans.  Odd ensemble sizes are a deliberate extension (the two
            # sets then differ by one walker).
            log.warning("nwalkers = %d < 2 * ndim = %d: the stretch move cannot leave the subspace the walkers span",
                        nwalkers, 2 * ndim)
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(ndim), float(a)
        self.use_graph = bool(use_graph)
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))   # follows np.random.seed like the reference's draws
        self.seed = int(seed)
        self._dc = device_chain
        # the self-check may rebuild the chain handle: it has to happen before the ensemble binds it
        if not device_chain._checked:
            device_chain._self_check()
        h = C.c_void_p()
        _lib.check(_lib.lib.gpbt_ensemble_create(C.byref(h), device_chain.handle(), self.nwalkers, self.a,
                                                 1 if randomize_split else 0, C.c_uint64(self.seed)))
        self._h = h
        device_chain._dependents.add(self)
        self._has_state = False

    # ---- lifetime --------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None:
            _lib.lib.gpbt_ensemble_destroy(self._h)
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

    # ---- state -----------------------------------------------------------------------------
    def set_state(self, coords, log_prob=None):
        x = np.ascontiguousarray(as_rows(coords, self.ndim))
        if x.shape[0] != self.nwalkers:
            raise ValueError("initial state has %d walkers, sampler has %d" % (x.shape[0], self.nwalkers))
        if not np.all(np.isfinite(x)):
            raise ValueError("the initial state contains non-finite coordinates")
        lp = None if log_prob is None else np.ascontiguousarray(log_prob, dtype=np.float64).reshape(self.nwalkers)
        _lib.check(_lib.lib.gpbt_ensemble_set_state(self._handle(), _lib.host_ptr(x), _lib.host_ptr(lp)))
        self._has_state = True

    def get_state(self):
        x, lp = np.empty((self.nwalkers, self.ndim)), np.empty(self.nwalkers)
        _lib.check(_lib.lib.gpbt_ensemble_get_state(self._handle(), _lib.host_ptr(x), _lib.host_ptr(lp)))
        return State(x, lp)

    def reset(self):
        _lib.check(_lib.lib.gpbt_ensemble_reset(self._handle()))

    # ---- running ---------------------------------------------------------------------------
    def advance(self, nsteps, u=None, partner=None, perm=None):
        """nsteps sampler steps on the device.  u / partner / perm are the test hooks of
        gpbt_ensemble_run (host-supplied random streams)."""
        if not self._has_state:
            raise RuntimeError("no initial state: call set_state or run_mcmc(initial_state, ...)")
        nh = (self.nwalkers + 1) // 2
        if u is not None:
            u = np.ascontiguousarray(u, dtype=np.float64).reshape(nsteps, 2, nh, 2)
            partner = np.ascontiguousarray(partner, dtype=np.int32).reshape(nsteps, 2, nh)
        if perm is not None:
            perm = np.ascontiguousarray(perm, dtype=np.int32).reshape(nsteps, self.nwalkers)
        _lib.check(_lib.lib.gpbt_ensemble_run(self._handle(), int(nsteps), _lib.host_ptr(u), _lib.host_ptr(partner),
                                              _lib.host_ptr(perm), 1 if self.use_graph else 0))

    def _start(self, initial_state, skip_initial_state_check):
        if initial_state is None:
            if not self._has_state:
                raise ValueError("cannot have initial_state=None before the sampler has run")
            return
        if isinstance(initial_state, State) or hasattr(initial_state, "coords"):
            coords, lp = initial_state.coords, getattr(initial_state, "log_prob", None)
        else:
            coords, lp = initial_state, None
        coords = as_rows(coords, self.ndim)
        if not skip_initial_state_check and self.nwalkers > 1:
            # emcee refuses walkers that do not span the space (linearly dependent start)
            c = np.atleast_2d(np.cov(coords, rowvar=False))
            d = np.sqrt(np.diag(c))
            if np.any(d == 0) or np.linalg.cond(c / np.outer(d, d)) > 1e8:
                raise ValueError("Initial state has a large condition number. Make sure that your walkers are "
                                 "linearly independent for the best performance")
        self.set_state(coords, lp)
        if np.any(np.isnan(self.get_state().log_prob)):
            raise ValueError("The initial log_prob was NaN")

    def sample(self, initial_state=None, iterations=1, skip_initial_state_check=False, chunk=None, **_):
        """Generator over the run in chunks of `chunk` device steps (default: all of them); yields
        the state after each chunk.  (emcee yields after every step; a step-by-step loop would put
        the host back into the loop this sampler removes.)"""
        self._start(initial_state, skip_initial_state_check)
        left = int(iterations)
        _lib.check(_lib.lib.gpbt_ensemble_reserve(self._handle(), left))
        chunk = left if not chunk else int(chunk)
        while left > 0:
            n = min(chunk, left)
            self.advance(n)
            left -= n
            yield self.get_state()

    def run_mcmc(self, initial_state, nsteps, status=None, skip_initial_state_check=False, **_):
        """LoggingEnsembleSampler.run_mcmc (src/mcmc.py:69-92): logs the acceptance fraction every
        `status` steps (default about 10 % of the run) and returns the final state."""
        log.info("running %d walkers for %d steps", self.nwalkers, nsteps)
        if not status:
            status = max(nsteps // 10, 1)
        state, done = None, 0
        for state in self.sample(initial_state, iterations=nsteps, chunk=status,
                                 skip_initial_state_check=skip_initial_state_check):
            done = min(done + status, nsteps)
            af = self.acceptance_fraction
            log.info("step %d: acceptance fraction: mean %.4f, std %.4f, min %.4f, max %.4f",
                     done, af.mean(), af.std(), af.min(), af.max())
        return state

    # ---- results ---------------------------------------------------------------------------
    @property
    def iteration(self):
        return int(_lib.lib.gpbt_ensemble_steps(self._handle()))

    def _read(self, want_chain, want_lp):
        n = self.iteration
        chain = np.empty((n, self.nwalkers, self.ndim)) if want_chain else None
        lp = np.empty((n, self.nwalkers)) if want_lp else None
        acc = np.zeros(self.nwalkers, dtype=np.int64)
        notpd = C.c_int64(0)
        _lib.check(_lib.lib.gpbt_ensemble_read(self._handle(), 0, n, _lib.host_ptr(chain), _lib.host_ptr(lp),
                                               _lib.host_ptr(acc), C.cast(C.byref(notpd), C.c_void_p)))
        self.n_notpd = notpd.value
        return chain, lp, acc

    def get_chain(self, flat=False, thin=1, discard=0):
        """[step, walker, dim] (emcee's layout); flat=True merges step and walker."""
        c = self._read(True, False)[0][discard::thin]
        return c.reshape(-1, self.ndim) if flat else c

    def get_log_prob(self, flat=False, thin=1, discard=0):
        lp = self._read(False, True)[1][discard::thin]
        return lp.reshape(-1) if flat else lp

    @property
    def n_accepted(self):
        """accepted proposals per walker since the last reset"""
        return self._read(False, False)[2]

    @property
    def acceptance_fraction(self):
        return self.n_accepted / max(self.iteration, 1)

    # emcee's backwards-compatible properties, which the reference reads (src/mcmc.py:383-388, 412)
    @property
    def chain(self):
        return np.swapaxes(self.get_chain(), 0, 1)          # [walker, step, dim]

    @property
    def lnprobability(self):
        return np.swapaxes(self.get_log_prob(), 0, 1)       # [walker, step]

    @property
    def flatchain(self):
        return self.get_chain(flat=True)

    @property
    def flatlnprobability(self):
        return self.get_log_prob(flat=True)


class ShardedEnsembleSampler(DeviceEnsembleSampler):
    """The same sampler over several GPUs, one process per GPU (torch.distributed): every rank keeps
    the whole ensemble and, with the same seed, makes identical proposals; rank r evaluates rows
    [r n_loc, (r + 1) n_loc) of each half-ensemble and the log-posteriors reach all ranks through the
    fused gather (dist.PeerGather: peer stores from the last kernel + one device barrier), after which
    every rank takes the same accept decisions.  The only data crossing NVLink is 8 bytes per proposal.
    Every rank must call the same methods with the same arguments; results are identical on all ranks."""

    def __init__(self, nwalkers, ndim, device_chain, seed, a=2.0, randomize_split=True, group=None):
        import torch
        from .dist import PeerGather
        if seed is None:
            raise ValueError("a sharded sampler needs an explicit seed (it must be the same on every rank)")
        super().__init__(nwalkers, ndim, device_chain, a=a, seed=seed, randomize_split=randomize_split,
                         use_graph=False)
        self._torch = torch
        self._dev_index = torch.cuda.current_device()
        nh = (self.nwalkers + 1) // 2
        import torch.distributed as dist
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_loc = -(-nh // self.world)
        self._pg = PeerGather(self.n_loc, torch.device("cuda", self._dev_index), group=group)
        self._x_loc = torch.empty((self.n_loc, self.ndim), dtype=torch.float64, device="cuda")

    def advance(self, nsteps, u=None, partner=None, perm=None):
        if u is not None or perm is not None:
            raise ValueError("host-supplied draws are a single-process test hook")
        if not self._has_state:
            raise RuntimeError("no initial state: call set_state or run_mcmc(initial_state, ...)")
        torch, lib, h = self._torch, _lib.lib, self._handle()
        _lib.check(lib.gpbt_ensemble_prepare(h, int(nsteps)))
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(int(nsteps)):
            for half in (0, 1):
                _lib.check(lib.gpbt_ensemble_begin_half(h, half, st))
                _lib.check(lib.gpbt_ensemble_copy_proposals(h, half, self.rank * self.n_loc, self.n_loc,
                                                            self._x_loc.data_ptr(), st))
                lp_all = self._pg.evaluate(self._dc, self._x_loc, -np.inf)
                _lib.check(lib.gpbt_ensemble_end_half(h, half, lp_all.data_ptr(), st))
        torch.cuda.current_stream().synchronize()
