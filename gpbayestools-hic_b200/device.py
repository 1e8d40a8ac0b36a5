"""
Device-side runners: thin Python over the C ABI (include/gpbt.h).  PyTorch is used only as the
device-memory allocator / stream provider for calls that hand tensors back to the caller.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from .state import EmulatorState

_COV_CHUNK_BYTES = 4 << 30  # device bytes of covariance produced per chunk by predict()
_FANOUT_MIN_ROWS = 256      # rows per GPU below which a further GPU does not pay (gpbt_fanout_*'s default)


def _require_gpu():
    if _lib.lib.gpbt_device_count() < 1:
        raise RuntimeError("gpbt_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("gpbt_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch


def _stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def as_rows(X, p=None):
    """float64 C-contiguous [N, p]; a 1-D input is one row (np.array(X, ndmin=2) in the reference)."""
    X = np.asarray(X, dtype=np.float64)   # no copy for float64 input
    if X.ndim < 2:
        X = X.reshape(1, -1)
    X = np.ascontiguousarray(X)
    if p is not None and X.shape[1] != p:
        raise ValueError("expected %d parameters per row, got %d" % (p, X.shape[1]))
    return X


class DeviceEmulator:
    """Emulator.predict on the GPU for one trained emulator (kernels (a) and (b))."""

    def __init__(self, state: EmulatorState):
        self.state = state

    # -- device tensors in, device tensors out ------------------------------------------------
    def pc_predict_device(self, X_d, extra_d=None):
        torch = _torch()
        st = self.state
        N = X_d.shape[0]
        zm = torch.empty((N, st.q), dtype=torch.float64, device=X_d.device)
        zv = torch.empty_like(zm)
        _lib.check(_lib.lib.gpbt_pc_predict(
            st.handle(), X_d.data_ptr(), None if extra_d is None else extra_d.data_ptr(),
            zm.data_ptr(), zv.data_ptr(), st.q, N, _stream_ptr(torch)))
        return zm, zv

    def predict_device(self, X_d, return_cov=True, extra_d=None):
        torch = _torch()
        st = self.state
        N = X_d.shape[0]
        zm, zv = self.pc_predict_device(X_d, extra_d)
        mean = torch.empty((N, st.m), dtype=torch.float64, device=X_d.device)
        cov = torch.empty((N, st.m, st.m), dtype=torch.float64, device=X_d.device) if return_cov else None
        _lib.check(_lib.lib.gpbt_backtransform(
            st.handle(), zm.data_ptr(), zv.data_ptr(), st.q, mean.data_ptr(), st.m,
            None if cov is None else cov.data_ptr(), st.m, 0, N, _stream_ptr(torch)))
        return (mean, cov) if return_cov else mean

    def predict_diag_device(self, X_d, extra_d=None):
        """mean [N, m] and diag(cov) [N, m] on the device (no N*m^2 covariance is formed)"""
        torch = _torch()
        st = self.state
        N = X_d.shape[0]
        zm, zv = self.pc_predict_device(X_d, extra_d)
        mean = torch.empty((N, st.m), dtype=torch.float64, device=X_d.device)
        var = torch.empty_like(mean)
        _lib.check(_lib.lib.gpbt_backtransform_diag(
            st.handle(), zm.data_ptr(), zv.data_ptr(), st.q, mean.data_ptr(), var.data_ptr(), st.m, 0, N,
            _stream_ptr(torch)))
        return mean, var

    def predict_diag(self, X, extra_std=0, chunk=1 << 18):
        """NumPy in/out: (mean, var) with var = diag of the covariance predict() would return."""
        torch = _torch()
        st = self.state
        X = as_rows(X, st.p_in)
        N = X.shape[0]
        extra = np.asarray(extra_std, dtype=np.float64).reshape(-1)
        if extra.size == 1:
            extra = None if extra[0] == 0.0 else np.full(N, extra[0])
        mean, var = np.empty((N, st.m)), np.empty((N, st.m))
        for s in range(0, N, chunk):
            e = min(N, s + chunk)
            X_d = torch.from_numpy(X[s:e]).cuda()
            extra_d = None if extra is None else torch.from_numpy(np.ascontiguousarray(extra[s:e])).cuda()
            m_d, v_d = self.predict_diag_device(X_d, extra_d)
            mean[s:e] = m_d.cpu().numpy()
            var[s:e] = v_d.cpu().numpy()
        return mean, var

    # -- numpy in, numpy out (the reference's signature) --------------------------------------
    def predict(self, X, return_cov=True, extra_std=0):
        torch = _torch()
        st = self.state
        X = as_rows(X, st.p_in)
        N = X.shape[0]
        extra = np.asarray(extra_std, dtype=np.float64).reshape(-1)
        if extra.size == 1:
            extra = None if extra[0] == 0.0 else np.full(N, extra[0])
        elif extra.size != N:
            raise ValueError("extra_std must be a scalar or have one entry per row of X")
        mean = np.empty((N, st.m))
        cov = np.empty((N, st.m, st.m)) if return_cov else None
        rows = N if not return_cov else max(1, min(N, _COV_CHUNK_BYTES // (8 * st.m * st.m)))
        for s in range(0, N, rows):
            e = min(N, s + rows)
            X_d = torch.from_numpy(X[s:e]).cuda()
            extra_d = None if extra is None else torch.from_numpy(np.ascontiguousarray(extra[s:e])).cuda()
            out = self.predict_device(X_d, return_cov, extra_d)
            if return_cov:
                mean[s:e] = out[0].cpu().numpy()
                cov[s:e] = out[1].cpu().numpy()
            else:
                mean[s:e] = out.cpu().numpy()
        return (mean, cov) if return_cov else mean


def mvn_loglike_batch(dY, cov, notpd_value=-np.inf, cov_add=None):
    """Batched mvn_loglike (src/mcmc.py:23-65) for dY [N, m], cov [N, m, m] (NumPy in/out); cov_add
    [m, m], if given, is added to every covariance on the device (cov + expdata_cov, src/mcmc.py:290)."""
    torch = _torch()
    dY = np.ascontiguousarray(np.array(dY, dtype=np.float64, ndmin=2))
    cov = np.ascontiguousarray(np.asarray(cov, dtype=np.float64)).reshape(dY.shape[0], dY.shape[1], dY.shape[1])
    N, m = dY.shape
    out = np.empty(N)
    add_d = None if cov_add is None else torch.from_numpy(
        np.ascontiguousarray(cov_add, dtype=np.float64).reshape(m, m)).cuda()
    rows = max(1, min(N, _COV_CHUNK_BYTES // (8 * m * m)))
    for s in range(0, N, rows):
        e = min(N, s + rows)
        y_d = torch.from_numpy(dY[s:e]).cuda()
        c_d = torch.from_numpy(cov[s:e]).cuda()   # device copy; the kernel factorises it in place
        lp_d = torch.empty(e - s, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib.gpbt_mvn_loglike(y_d.data_ptr(), None, c_d.data_ptr(),
                                             None if add_d is None else add_d.data_ptr(), lp_d.data_ptr(),
                                             None, float(notpd_value), e - s, m, _stream_ptr(torch)))
        out[s:e] = lp_d.cpu().numpy()
    return out


def default_devices():
    """CUDA devices a single-process Chain fans large host batches out over.
    GPBT_DEVICES = "all" | "1" (the current device only) | "0,2,3" (explicit list, first = primary).
    Default: every visible device -- unless this process is one rank of a multi-process job
    (WORLD_SIZE > 1, torchrun), where each rank keeps to its own GPU."""
    import os
    cur = _lib.lib.gpbt_get_device()
    spec = os.environ.get("GPBT_DEVICES", "").strip().lower()
    if spec in ("1", "one", "current") or int(os.environ.get("WORLD_SIZE", "1") or 1) > 1:
        return [cur]
    n = _lib.lib.gpbt_device_count()
    if spec and spec != "all":
        devs = [int(t) for t in spec.split(",") if t.strip() != ""]
        bad = [d for d in devs if d < 0 or d >= n]
        if bad or not devs:
            raise ValueError("GPBT_DEVICES=%r names devices outside 0..%d" % (spec, n - 1))
        return devs
    return [cur] + [d for d in range(n) if d != cur]


class DeviceChain:
    """Everything between X[N, p] and lp[N] for a list of emulators + experimental data
    (gpbt_chain_t).  Builds the low-rank factors on the host when every emulator is PCA-mode.

    device   CUDA device of this chain (default: the one current at first use).
    devices  devices the HOST call (log_target) may fan a large batch out over, this chain's own device
             first; replicas on the others are created the first time a batch is large enough
             (gpbt_fanout_*: one worker thread and one pinned staging buffer per GPU, no collective)."""

    def __init__(self, states, lo, hi, y_exp, cov_exp, lowrank=True, device=None, devices=None):
        self.states = list(states)
        self.p = self.states[0].p_in   # columns of X (before any parameter-function pre-transform)
        self.M = sum(s.m for s in self.states)
        self.Q = sum(s.q for s in self.states)
        self.lo = np.ascontiguousarray(lo, dtype=np.float64).reshape(self.p)
        self.hi = np.ascontiguousarray(hi, dtype=np.float64).reshape(self.p)
        self.y_exp = np.ascontiguousarray(y_exp, dtype=np.float64).reshape(self.M)
        self.cov_exp = np.ascontiguousarray(cov_exp, dtype=np.float64).reshape(self.M, self.M)
        self.device = device
        self.devices = None if devices is None else list(devices)
        if self.devices:
            if device is not None and self.devices[0] != device:
                raise ValueError("devices[0] must be the chain's own device")
            self.device = self.devices[0]
        self.lowrank = None
        self.lowrank_error = None
        if lowrank is not False and lowrank is not None and not isinstance(lowrank, bool):
            self.lowrank = lowrank          # factors handed over by the chain this one replicates
        elif lowrank and all(not (s.no_pca or s.exp_diag) for s in self.states) and self.Q <= self.M:
            # F = blockdiag(Ctrunc) + cov_exp may be singular although every C_w is positive definite
            # (zero truncation term of a PCGP emulator + an experimental error of exactly 0): the dense
            # path handles that case, the factorisation of F cannot
            try:
                lr = lowrank_factors(self.states, self.y_exp, self.cov_exp)
                vals = [lr["s_perp"], lr["logdetF_half"], float(np.abs(lr["R"]).max()), float(np.abs(lr["c0"]).max())]
                if not np.all(np.isfinite(vals)):
                    raise np.linalg.LinAlgError("non-finite low-rank factors")
                self.lowrank = lr
            except (np.linalg.LinAlgError, ValueError) as exc:
                import warnings
                self.lowrank_error = str(exc)
                warnings.warn("gpbt_b200: truncation + experimental covariance cannot be factorised (%s); "
                              "using the dense path" % exc)
        self._handle = None
        self._checked = self.lowrank is None
        self.lowrank_check = None
        self._dependents = weakref.WeakSet()   # device samplers bound to the handle
        self._replicas = []                    # DeviceChain per further device (fan-out)
        self._fanout = None
        self.last_devices_used = 1

    def handle(self):
        if self._handle is None:
            if self.device is None:
                self.device = _lib.lib.gpbt_get_device()
            h = C.c_void_p()
            hp = _lib.host_ptr
            lr = self.lowrank
            with _lib.on_device(self.device):
                arr = (C.c_void_p * len(self.states))(*[s.handle(self.device) for s in self.states])
                _lib.check(_lib.lib.gpbt_chain_create(
                    C.byref(h), arr, len(self.states), self.p, hp(self.lo), hp(self.hi), hp(self.y_exp),
                    hp(self.cov_exp), hp(lr["R"]) if lr else None, hp(lr["c0"]) if lr else None,
                    lr["s_perp"] if lr else 0.0, lr["logdetF_half"] if lr else 0.0))
            self._handle = h
        return self._handle

    def fanout(self):
        """gpbt_fanout_t over this chain and one replica per further device of `devices`"""
        if self._fanout is None:
            own = self.handle()
            self._replicas = [DeviceChain(self.states, self.lo, self.hi, self.y_exp, self.cov_exp,
                                          lowrank=self.lowrank if self.lowrank is not None else False, device=d)
                              for d in self.devices[1:]]
            for r in self._replicas:
                r._checked = True
            handles = [own] + [r.handle() for r in self._replicas]
            f = C.c_void_p()
            _lib.check(_lib.lib.gpbt_fanout_create(C.byref(f), (C.c_void_p * len(handles))(*handles), len(handles)))
            self._fanout = f
        return self._fanout

    def _self_check(self, n_points=16, tol=1e-8):
        """Once per chain: the exact low-rank path against the dense Cholesky path on a few points of
        the box.  The identity is exact, but its host-side factors go through chol(F) and a QR; if F
        (truncation + experimental covariance) is so ill-conditioned that the two paths drift apart,
        the chain is rebuilt without the low-rank factors and every call takes the dense path."""
        self._checked = True
        rng = np.random.default_rng(20261018)
        X = rng.uniform(self.lo, self.hi, (n_points, self.p))
        a = self._call_host(X, -np.inf, _lib.PATH_LOWRANK, 1)
        b = self._call_host(X, -np.inf, _lib.PATH_DENSE, 1)
        ok = np.isfinite(a) & np.isfinite(b)
        diff = float(np.max(np.abs(a[ok] - b[ok]))) if ok.any() else 0.0
        self.lowrank_check = dict(points=int(ok.sum()), max_abs_diff=diff, tol=tol)
        if not np.array_equal(np.isfinite(a), np.isfinite(b)) or diff > tol * max(1.0, float(np.max(np.abs(b[ok]), initial=0.0)) / 100.0):
            import warnings
            warnings.warn("gpbt_b200: low-rank and dense log-likelihood paths differ by %.2e on this chain "
                          "(ill-conditioned truncation + experimental covariance?); using the dense path" % diff)
            self.release()
            self.lowrank = None

    def _call_host(self, X, oob_value, path, max_devices=None):
        lp = np.empty(X.shape[0])
        notpd = C.c_int(0)
        N = X.shape[0]
        if self.devices and len(self.devices) > 1 and max_devices != 1 and (max_devices or N >= 2 * _FANOUT_MIN_ROWS):
            used = C.c_int(0)
            _lib.check(_lib.lib.gpbt_fanout_log_posterior_host(
                self.fanout(), _lib.host_ptr(X), float(oob_value), _lib.host_ptr(lp), C.byref(notpd), N, path,
                int(max_devices or 0), C.byref(used)))
            self.last_devices_used = used.value
        else:
            with _lib.on_device(self.device):
                _lib.check(_lib.lib.gpbt_log_posterior_host(
                    self.handle(), _lib.host_ptr(X), float(oob_value), _lib.host_ptr(lp), C.byref(notpd), N, path))
            self.last_devices_used = 1
        self.last_notpd = notpd.value
        return lp

    def release(self):
        for dep in list(getattr(self, "_dependents", ())):
            dep.close()
        if getattr(self, "_fanout", None) is not None:
            _lib.lib.gpbt_fanout_destroy(self._fanout)   # joins the worker threads
            self._fanout = None
        for r in getattr(self, "_replicas", ()):
            r.release()
        self._replicas = []
        if self._handle is not None:
            _lib.lib.gpbt_chain_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass   # interpreter shutdown

    @staticmethod
    def _path(path):
        return {None: _lib.PATH_AUTO, "auto": _lib.PATH_AUTO, "dense": _lib.PATH_DENSE,
                "lowrank": _lib.PATH_LOWRANK, "diag": _lib.PATH_DIAG}[path]

    def log_target(self, X, oob_value, path=None, max_devices=None):
        """Host buffers in/out: gpbt_log_posterior_host (H2D, kernels, D2H, one sync) on this chain's GPU,
        or -- with `devices` and a batch of at least 2 x 256 rows -- gpbt_fanout_log_posterior_host over
        several GPUs.  max_devices: None = automatic, 1 = this GPU only, n = use n GPUs whatever N is."""
        _require_gpu()
        X = as_rows(X, self.p)
        if not self._checked:
            self._self_check()
        return self._call_host(X, oob_value, self._path(path), max_devices)

    def log_target_device(self, X_d, oob_value, lp_d=None, path=None):
        """Device tensors in/out on torch's current stream; no synchronisation."""
        torch = _torch()
        if not self._checked:
            self._self_check()
        N = X_d.shape[0]
        if lp_d is None:
            lp_d = torch.empty(N, dtype=torch.float64, device=X_d.device)
        _lib.check(_lib.lib.gpbt_log_posterior(
            self.handle(), X_d.data_ptr(), float(oob_value), lp_d.data_ptr(), None, N,
            self._path(path), _stream_ptr(torch)))
        return lp_d

    def log_target_scatter(self, X_d, oob_value, peer_ptrs, peer_off, lp_d=None, path=None):
        """log_target_device with the all-gather fused in: each result is also stored to
        peer_ptrs[r] + peer_off (device pointers of peer-mapped buffers).  No synchronisation."""
        torch = _torch()
        if not self._checked:
            self._self_check()
        N = X_d.shape[0]
        if lp_d is None:
            lp_d = torch.empty(N, dtype=torch.float64, device=X_d.device)
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        _lib.check(_lib.lib.gpbt_log_posterior_scatter(
            self.handle(), X_d.data_ptr(), float(oob_value), lp_d.data_ptr(), C.cast(arr, C.c_void_p),
            len(peer_ptrs), int(peer_off), None, N, self._path(path), _stream_ptr(torch)))
        return lp_d

    def predict(self, X, extra_std=0.0, return_cov=True):
        """Chain._predict (src/mcmc.py:153-166): mean [N, M], block-diagonal cov [N, M, M]."""
        torch = _torch()
        X = as_rows(X, self.p)
        N = X.shape[0]
        mean = np.empty((N, self.M))
        cov = np.empty((N, self.M, self.M)) if return_cov else None
        rows = N if not return_cov else max(1, min(N, _COV_CHUNK_BYTES // (8 * self.M * self.M)))
        for s in range(0, N, rows):
            e = min(N, s + rows)
            X_d = torch.from_numpy(X[s:e]).cuda()
            mean_d = torch.empty((e - s, self.M), dtype=torch.float64, device="cuda")
            cov_d = torch.empty((e - s, self.M, self.M), dtype=torch.float64, device="cuda") if return_cov else None
            _lib.check(_lib.lib.gpbt_chain_predict(
                self.handle(), X_d.data_ptr(), float(extra_std), mean_d.data_ptr(),
                None if cov_d is None else cov_d.data_ptr(), e - s, _stream_ptr(torch)))
            mean[s:e] = mean_d.cpu().numpy()
            if return_cov:
                cov[s:e] = cov_d.cpu().numpy()
        return (mean, cov) if return_cov else mean


def lowrank_factors(states, y_exp, cov_exp):
    """Host set-up (once per chain) of the exact low-rank form used by lowrank_loglike.cuh:
        C_w = F + U^T diag(v_w) U,  F = blockdiag(Ctrunc_e) + cov_exp,  U = blockdiag(A_e)
        L_F^-1 U^T = Qb R,  c0 = Qb^T L_F^-1 (mu - y_exp),  s_perp = |(I - Qb Qb^T) L_F^-1 (mu - y_exp)|^2
    (reference quantities: src/emulator.py:335-363, src/mcmc.py:153-166, 288-290).
    When cov_exp does not couple the emulators (block diagonal, e.g. the diagonal matrix the reference
    reads from its pickle) the factorisation is done per emulator: R comes out exactly block diagonal
    and the device evaluates log L as a sum over emulator blocks."""
    M = sum(s.m for s in states)
    Q = sum(s.q for s in states)
    cov_exp = np.asarray(cov_exp, dtype=np.float64)
    mask = np.zeros((M, M), dtype=bool)
    mo = 0
    for s in states:
        mask[mo:mo + s.m, mo:mo + s.m] = True
        mo += s.m
    if len(states) > 1 and not np.any(cov_exp[~mask]):
        R, c0, s_perp, ld = np.zeros((Q, Q)), np.empty(Q), 0.0, 0.0
        qo = mo = 0
        for s in states:
            part = _lowrank_factors_joint([s], y_exp[mo:mo + s.m], cov_exp[mo:mo + s.m, mo:mo + s.m])
            R[qo:qo + s.q, qo:qo + s.q] = part["R"]
            c0[qo:qo + s.q] = part["c0"]
            s_perp += part["s_perp"]
            ld += part["logdetF_half"]
            qo += s.q
            mo += s.m
        return dict(R=np.ascontiguousarray(R), c0=np.ascontiguousarray(c0), s_perp=float(s_perp),
                    logdetF_half=float(ld))
    return _lowrank_factors_joint(states, y_exp, cov_exp)


def _lowrank_factors_joint(states, y_exp, cov_exp):
    from scipy.linalg import cholesky, qr, solve_triangular
    M = sum(s.m for s in states)
    Q = sum(s.q for s in states)
    F = np.array(cov_exp, dtype=np.float64)
    U = np.zeros((Q, M))
    mu = np.empty(M)
    qo = mo = 0
    for s in states:
        F[mo:mo + s.m, mo:mo + s.m] += s.Ctrunc
        U[qo:qo + s.q, mo:mo + s.m] = s.A
        mu[mo:mo + s.m] = s.mu
        qo += s.q
        mo += s.m
    LF = cholesky(F, lower=True, check_finite=True)
    Bt = solve_triangular(LF, U.T, lower=True)          # [M, Q]
    Qb, R = qr(Bt, mode="economic")                     # Bt = Qb R
    sgn = np.sign(np.diag(R))
    sgn[sgn == 0] = 1.0
    Qb, R = Qb * sgn, R * sgn[:, None]                  # positive diagonal (cosmetic)
    w0 = solve_triangular(LF, mu - y_exp, lower=True)
    c0 = Qb.T @ w0
    perp = w0 - Qb @ c0
    perp -= Qb @ (Qb.T @ perp)                          # one re-orthogonalisation step
    return dict(R=np.ascontiguousarray(np.triu(R)), c0=np.ascontiguousarray(c0),
                s_perp=float(perp @ perp), logdetF_half=float(np.log(np.diag(LF)).sum()))
