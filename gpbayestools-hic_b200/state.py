"""
Trained-emulator state: the hand-off between (offline, CPU, scikit-learn) training and the
device-resident hot path.

The reference keeps its trained state on sklearn objects inside the dill-ed Emulator
(src/emulator.py:309-363; loaded by Chain.loadEmulator, src/mcmc.py:145-150).  `EmulatorState`
pulls out exactly the arrays the prediction path reads and uploads them through
gpbt_emulator_create (include/gpbt.h).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
from scipy.linalg import lapack

GPR_ALPHA = 0.1  # GaussianProcessRegressor(alpha=0.1) in the reference (src/emulator.py:310)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def invert_lower(L):
    """Explicit inverse of a lower-triangular Cholesky factor (LAPACK dtrtri, FP64).
    SURVEY 7(ii): the triangular form W = L^-1 keeps |W k|^2 as accurate as the triangular solve,
    whereas K^-1 = W^T W as a quadratic form loses 3 digits."""
    W, info = lapack.dtrtri(np.asarray(L, dtype=np.float64, order="F"), lower=1)
    if info != 0:
        raise np.linalg.LinAlgError("dtrtri failed: info=%d" % info)
    return np.tril(W)


def cholesky_of_kernel(kind, Xtr, c, ell, sn):
    """L_ as sklearn's fit builds it: cholesky(kernel_(X_train) + alpha*I) (_gpr.py:349-360)."""
    from scipy.linalg import cholesky
    from scipy.spatial.distance import pdist, squareform
    Xs = Xtr / ell
    if kind == "RBF":
        K = squareform(np.exp(-0.5 * pdist(Xs, metric="sqeuclidean")))
        np.fill_diagonal(K, 1.0)
    elif kind == "Matern":
        r = squareform(pdist(Xs, metric="euclidean")) * np.sqrt(3.0)
        K = (1.0 + r) * np.exp(-r)
    else:
        raise ValueError("unknown kernel kind %r" % (kind,))
    K = c * K
    K[np.diag_indices_from(K)] += sn
    K[np.diag_indices_from(K)] += GPR_ALPHA
    return cholesky(K, lower=True, check_finite=False)


def _kernel_kind_and_hypers(kernel):
    """(kind, c, ell, sn) of a fitted `C * {RBF|Matern(nu=1.5)} + White` kernel
    (the only form the reference trains, src/emulator.py:288-307)."""
    from sklearn.gaussian_process import kernels as sk
    try:
        prod, white = kernel.k1, kernel.k2
        const, base = prod.k1, prod.k2
    except AttributeError as e:
        raise TypeError("expected (Constant * RBF|Matern) + White, got %r" % (kernel,)) from e
    if not isinstance(white, sk.WhiteKernel) or not isinstance(const, sk.ConstantKernel):
        raise TypeError("expected (Constant * RBF|Matern) + White, got %r" % (kernel,))
    if isinstance(base, sk.Matern):
        if base.nu != 1.5:
            raise NotImplementedError("only Matern(nu=1.5) is on the reference's path")
        kind = "Matern"
    elif isinstance(base, sk.RBF):
        kind = "RBF"
    else:
        raise TypeError("unsupported base kernel %r" % (base,))
    return kind, float(const.constant_value), np.asarray(base.length_scale, dtype=np.float64), \
        float(white.noise_level)


@dataclass
class EmulatorState:
    kind: str                 # "RBF" | "Matern"
    Xtr: np.ndarray           # [n, p]
    ell: np.ndarray           # [q, p]
    c: np.ndarray             # [q]
    sn: np.ndarray            # [q]
    alpha: np.ndarray         # [q, n]
    Linv: np.ndarray          # [q, n, n] lower-triangular inverse of L_
    mu: np.ndarray            # [m]
    scale: np.ndarray         # [m]
    A: Optional[np.ndarray] = None        # [q, m]   (PCA mode)
    Ctrunc: Optional[np.ndarray] = None   # [m, m]   (PCA mode)
    no_pca: bool = False
    exp_diag: bool = False
    L: Optional[np.ndarray] = None        # [q, n, n] kept only for oracle comparisons
    _handle: Optional[C.c_void_p] = field(default=None, repr=False, compare=False)

    # ---- shapes ---------------------------------------------------------------------------
    @property
    def p(self): return self.Xtr.shape[1]
    @property
    def n(self): return self.Xtr.shape[0]
    @property
    def q(self): return self.alpha.shape[0]
    @property
    def m(self): return self.mu.shape[0]

    # ---- constructors ---------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, kind, Xtr, ell, c, sn, alpha, mu, scale, A=None, Ctrunc=None, L=None,
                    no_pca=False, exp_diag=False, keep_L=True):
        Xtr, alpha = _f64(Xtr), _f64(alpha)
        q, n = alpha.shape
        ell = _f64(np.broadcast_to(np.asarray(ell, dtype=np.float64).reshape(q, -1), (q, Xtr.shape[1])))
        c, sn = _f64(c).reshape(q), _f64(sn).reshape(q)
        if L is None:
            L = np.stack([cholesky_of_kernel(kind, Xtr, c[j], ell[j], sn[j]) for j in range(q)])
        L = _f64(L)
        Linv = np.stack([invert_lower(L[j]) for j in range(q)])
        return cls(kind=str(kind), Xtr=Xtr, ell=ell, c=c, sn=sn, alpha=alpha, Linv=_f64(Linv),
                   mu=_f64(mu), scale=_f64(scale), A=None if A is None else _f64(A),
                   Ctrunc=None if Ctrunc is None else _f64(Ctrunc), no_pca=bool(no_pca),
                   exp_diag=bool(exp_diag), L=L if keep_L else None)

    @classmethod
    def from_trained(cls, emu, keep_L=False):
        """From any trained emulator object exposing the reference's attributes: `gps` (sklearn
        GPRs), `scaler`, `npc`, `_trans_matrix`, `_cov_trunc`, `perform_no_PCA_`,
        `exp_and_cov_diagonal_` -- i.e. a dill-loaded reference `src.emulator.Emulator` or this
        package's `Emulator`."""
        if getattr(emu, "parameterTrafoPCA_", False):
            raise NotImplementedError(
                "parameterTrafoPCA emulators need the host pre-transform (src/emulator.py:492-551); "
                "not on the accelerated path yet")
        gps = emu.gps
        hyp = [_kernel_kind_and_hypers(g.kernel_) for g in gps]
        kinds = {h[0] for h in hyp}
        if len(kinds) != 1:
            raise ValueError("mixed kernel families in one emulator")
        Xtr = _f64(gps[0].X_train_)
        p = Xtr.shape[1]
        no_pca = bool(getattr(emu, "perform_no_PCA_", False))
        return cls.from_arrays(
            kind=kinds.pop(), Xtr=Xtr,
            ell=np.stack([np.broadcast_to(h[2], (p,)) for h in hyp]),
            c=[h[1] for h in hyp], sn=[h[3] for h in hyp],
            alpha=np.stack([np.asarray(g.alpha_, dtype=np.float64).reshape(-1) for g in gps]),
            mu=emu.scaler.mean_, scale=emu.scaler.scale_,
            A=None if no_pca else emu._trans_matrix[:emu.npc],
            Ctrunc=None if no_pca else emu._cov_trunc,
            L=np.stack([g.L_ for g in gps]), no_pca=no_pca,
            exp_diag=bool(getattr(emu, "exp_and_cov_diagonal_", False)), keep_L=keep_L)

    # ---- views ----------------------------------------------------------------------------
    def oracle_dict(self):
        """The layout oracle/gp_oracle.py works on (tests only)."""
        if self.L is None:
            raise ValueError("state was built with keep_L=False")
        d = dict(kind=self.kind, Xtr=self.Xtr, ell=self.ell, c=self.c, sn=self.sn, alpha=self.alpha,
                 L=self.L, no_pca=self.no_pca, exp_diag=self.exp_diag, mu=self.mu, scale=self.scale)
        if not self.no_pca:
            d.update(A=self.A, Ctrunc=self.Ctrunc)
        return d

    def device_bytes(self):
        n_pad = -(-self.n // 32) * 32
        return 8 * (self.q * n_pad * n_pad + self.q * n_pad * (self.p + 3) + self.m * self.m + self.q * self.m)

    # ---- device ---------------------------------------------------------------------------
    def handle(self):
        """gpbt_emulator_t for the current CUDA device (created on first use)."""
        if self._handle is None:
            from . import _lib
            h = C.c_void_p()
            flags = (_lib.FLAG_NO_PCA if self.no_pca else 0) | (_lib.FLAG_EXP_DIAG if self.exp_diag else 0)
            kind = {"RBF": _lib.KERNEL_RBF, "Matern": _lib.KERNEL_MATERN32}[self.kind]
            hp = _lib.host_ptr
            _lib.check(_lib.lib.gpbt_emulator_create(
                C.byref(h), self.p, self.n, self.q, self.m, kind, flags, hp(self.Xtr), hp(self.ell),
                hp(self.c), hp(self.sn), hp(self.alpha), hp(self.Linv), hp(self.A), hp(self.mu),
                hp(self.scale), hp(self.Ctrunc)))
            self._handle = h
        return self._handle

    def release(self):
        if self._handle is not None:
            from . import _lib
            _lib.lib.gpbt_emulator_destroy(self._handle)
            self._handle = None

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_handle"] = None
        return d
