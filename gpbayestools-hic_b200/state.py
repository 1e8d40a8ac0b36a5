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


# (kind, attribute suffix on the reference object, index attribute, grid) of the three parameter
# groups of "parameterTrafoPCA" in the order the reference applies them (src/emulator.py:79-99, 492-551)
PARAM_TRAFO_GROUPS = (
    (0, "bulk", "indices_zeta_s_parameters", (0.0, 0.5, 100)),
    (1, "shear", "indices_eta_s_parameters", (0.0, 0.6, 100)),
    (2, "yloss", "indices_yloss_parameters", (0.0, 6.2, 100)),
)


@dataclass
class ParamTrafo:
    """The parameter-function PCA pre-transform of one emulator, as plain arrays.
    groups[i] = dict(kind, idx [nargs], grid (lo, hi, npts), smean [npts], sscale [npts],
                     pmean [npts], comp [ncomp, npts])"""
    p_in: int
    groups: list

    @property
    def keep(self):
        used = {int(c) for g in self.groups for c in g["idx"]}
        return np.array([c for c in range(self.p_in) if c not in used], dtype=np.int32)

    @property
    def p_out(self):
        return len(self.keep) + sum(g["comp"].shape[0] for g in self.groups)

    @classmethod
    def from_trained(cls, emu):
        groups = []
        for kind, tag, idx_attr, grid in PARAM_TRAFO_GROUPS:
            sc, pc = getattr(emu, "paramTrafoScaler_" + tag), getattr(emu, "paramTrafoPCA_" + tag)
            groups.append(dict(kind=kind, idx=np.asarray(getattr(emu, idx_attr), dtype=np.int32), grid=grid,
                               smean=_f64(sc.mean_), sscale=_f64(sc.scale_), pmean=_f64(pc.mean_),
                               comp=_f64(pc.components_)))
        return cls(p_in=int(np.asarray(emu.design_points_org_).shape[1]), groups=groups)

    def folded(self):
        """(Wt [ncomp, npts], b [ncomp]) per group:  PC = f @ Wt.T + b  ==  PCA.transform(Scaler.transform(f))"""
        out = []
        for g in self.groups:
            Wt = np.ascontiguousarray(g["comp"] / g["sscale"])
            b = -(g["comp"] @ (g["smean"] / g["sscale"] + g["pmean"]))
            out.append((Wt, np.ascontiguousarray(b)))
        return out


@dataclass
class EmulatorState:
    kind: str                 # "RBF" | "Matern" | "PCGP" (surmise PCGP / PCSK, see from_pcgp_fitinfo)
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
    trafo: Optional[ParamTrafo] = None    # parameterTrafoPCA pre-transform (then p counts transformed columns)
    sig2: Optional[np.ndarray] = None     # [q]  PCGP only; there c / sn hold (1-nug)a / (1-nug)b, alpha = pw, Linv = Vh^T
    pcgp: Optional[dict] = None           # PCGP only: hypcov [q,p+1], nug [q], Vh [q,n,n] (oracle comparisons)
    _handles: dict = field(default_factory=dict, repr=False, compare=False)   # CUDA device index -> gpbt_emulator_t

    # ---- shapes ---------------------------------------------------------------------------
    @property
    def p(self): return self.Xtr.shape[1]
    @property
    def p_in(self): return self.trafo.p_in if self.trafo is not None else self.Xtr.shape[1]
    @property
    def n(self): return self.Xtr.shape[0]
    @property
    def q(self): return self.alpha.shape[0]
    @property
    def m(self): return self.mu.shape[0]

    # ---- constructors ---------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, kind, Xtr, ell, c, sn, alpha, mu, scale, A=None, Ctrunc=None, L=None,
                    no_pca=False, exp_diag=False, keep_L=True, trafo=None):
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
                   exp_diag=bool(exp_diag), L=L if keep_L else None, trafo=trafo)

    @classmethod
    def from_trained(cls, emu, keep_L=False):
        """From any trained emulator object exposing the reference's attributes: `gps` (sklearn
        GPRs), `scaler`, `npc`, `_trans_matrix`, `_cov_trunc`, `perform_no_PCA_`,
        `exp_and_cov_diagonal_` -- i.e. a dill-loaded reference `src.emulator.Emulator` or this
        package's `Emulator`."""
        trafo = ParamTrafo.from_trained(emu) if getattr(emu, "parameterTrafoPCA_", False) else None
        if not hasattr(emu, "gps") and hasattr(emu, "emu"):
            # EmulatorBAND: the trained object is a surmise emulator (src/emulator_BAND.py:270-292)
            info = emu.emu if isinstance(emu.emu, dict) else emu.emu._info
            return cls.from_pcgp_fitinfo(info, exp_diag=bool(getattr(emu, "exp_and_cov_diagonal_", False)), trafo=trafo)
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
            exp_diag=bool(getattr(emu, "exp_and_cov_diagonal_", False)), keep_L=keep_L, trafo=trafo)

    @classmethod
    def from_pcgp_fitinfo(cls, info, exp_diag=False, extravar_in_cov=False, trafo=None):
        """From the fit information of a surmise PCGP / PCSK emulator (`emulator._info` of the object
        the reference keeps in EmulatorBAND.emu, src/emulator_BAND.py:270-292):
            theta [n,p], pct [m,q], scale [m], offset [m], extravar [m],
            emulist[k] = {hypcov [p+1], hypind, nug, Vh [n,n], pw [n], sig2}
        PCs that share hyper-parameters (hypind) are expanded: the kernels evaluate every PC anyway.
        covx() of surmise 0.2.1 is pct diag(predvar) pct^T without extravar as far as its published
        source shows; `extravar_in_cov=True` adds diag(extravar).  PARITY UNPINNED (no surmise here)."""
        theta = _f64(info["theta"])
        emul = info["emulist"]
        q, (n, p) = len(emul), theta.shape
        hyp = np.stack([_f64(emul[int(e.get("hypind", k))]["hypcov"]).reshape(p + 1) for k, e in enumerate(emul)])
        nug = np.array([float(np.squeeze(e["nug"])) for e in emul])
        g = np.exp(hyp[:, -1])
        pct = _f64(info["pct"] if "pct" in info else info["pcti"]).reshape(-1, q)
        scale = _f64(info["scale"]).reshape(-1)
        m = scale.shape[0]
        Vh = np.stack([_f64(e["Vh"]).reshape(n, n) for e in emul])
        extra = np.diag(_f64(info["extravar"]).reshape(m)) if extravar_in_cov else np.zeros((m, m))
        return cls(kind="PCGP", Xtr=theta, ell=_f64(np.exp(hyp[:, :-1])), c=_f64((1.0 - nug) / (1.0 + g)),
                   sn=_f64((1.0 - nug) * g / (1.0 + g)),
                   alpha=np.stack([_f64(e["pw"]).reshape(n) for e in emul]),
                   Linv=_f64(np.swapaxes(Vh, 1, 2)), mu=_f64(info["offset"]).reshape(m), scale=np.ones(m),
                   A=_f64((pct * scale[:, None]).T), Ctrunc=_f64(extra), no_pca=False, exp_diag=bool(exp_diag),
                   L=None, trafo=trafo, sig2=_f64([float(np.squeeze(e["sig2"])) for e in emul]),
                   pcgp=dict(hypcov=hyp, nug=nug, Vh=Vh))

    # ---- views ----------------------------------------------------------------------------
    def oracle_dict(self):
        """The layout oracle/gp_oracle.py works on (tests only)."""
        if self.kind == "PCGP":
            d = dict(kind="PCGP", Xtr=self.Xtr, pw=self.alpha, sig2=self.sig2, no_pca=False, exp_diag=self.exp_diag,
                     mu=self.mu, scale=self.scale, A=self.A, Ctrunc=self.Ctrunc, **self.pcgp)
            if self.trafo is not None:
                d["trafo"] = dict(p_in=self.trafo.p_in, groups=self.trafo.groups)
            return d
        if self.L is None:
            raise ValueError("state was built with keep_L=False")
        d = dict(kind=self.kind, Xtr=self.Xtr, ell=self.ell, c=self.c, sn=self.sn, alpha=self.alpha,
                 L=self.L, no_pca=self.no_pca, exp_diag=self.exp_diag, mu=self.mu, scale=self.scale)
        if not self.no_pca:
            d.update(A=self.A, Ctrunc=self.Ctrunc)
        if self.trafo is not None:
            d["trafo"] = dict(p_in=self.trafo.p_in, groups=self.trafo.groups)
        return d

    def device_bytes(self):
        n_pad = -(-self.n // 32) * 32
        return 8 * (self.q * n_pad * n_pad + self.q * n_pad * (self.p + 3) + self.m * self.m + self.q * self.m)

    # ---- device ---------------------------------------------------------------------------
    def handle(self, device=None):
        """gpbt_emulator_t on CUDA device `device` (default: the current one), created on first use;
        one replica of the trained state per device."""
        from . import _lib
        if device is None:
            device = _lib.lib.gpbt_get_device()
        h = self._handles.get(device)
        if h is None:
            with _lib.on_device(device):
                h = self._create_handle(_lib)
            self._handles[device] = h
        return h

    def _create_handle(self, _lib):
        h = C.c_void_p()
        flags = (_lib.FLAG_NO_PCA if self.no_pca else 0) | (_lib.FLAG_EXP_DIAG if self.exp_diag else 0)
        hp = _lib.host_ptr
        if self.kind == "PCGP":
            _lib.check(_lib.lib.gpbt_emulator_create_pcgp(
                C.byref(h), self.p, self.n, self.q, self.m, flags, hp(self.Xtr), hp(self.ell), hp(self.c),
                hp(self.sn), hp(self.sig2), hp(self.alpha), hp(self.Linv), hp(self.A), hp(self.mu),
                hp(self.Ctrunc)))
        else:
            kind = {"RBF": _lib.KERNEL_RBF, "Matern": _lib.KERNEL_MATERN32}[self.kind]
            _lib.check(_lib.lib.gpbt_emulator_create(
                C.byref(h), self.p, self.n, self.q, self.m, kind, flags, hp(self.Xtr), hp(self.ell),
                hp(self.c), hp(self.sn), hp(self.alpha), hp(self.Linv), hp(self.A), hp(self.mu),
                hp(self.scale), hp(self.Ctrunc)))
        if self.trafo is not None:
            t = self.trafo
            ng = len(t.groups)
            keep = t.keep
            kinds = np.array([g["kind"] for g in t.groups], dtype=np.int32)
            idx = np.zeros((ng, 4), dtype=np.int32)
            for i, g in enumerate(t.groups):
                idx[i, :len(g["idx"])] = g["idx"]
            ncomp = np.array([g["comp"].shape[0] for g in t.groups], dtype=np.int32)
            npts = np.array([g["grid"][2] for g in t.groups], dtype=np.int32)
            glo = np.array([g["grid"][0] for g in t.groups], dtype=np.float64)
            ghi = np.array([g["grid"][1] for g in t.groups], dtype=np.float64)
            folded = t.folded()
            Wt = (C.c_void_p * ng)(*[hp(w) for w, _ in folded])
            bb = (C.c_void_p * ng)(*[hp(b) for _, b in folded])
            _lib.check(_lib.lib.gpbt_emulator_set_param_trafo(
                h, t.p_in, hp(keep), len(keep), ng, hp(kinds), hp(idx), hp(ncomp), hp(npts), hp(glo), hp(ghi),
                C.cast(Wt, C.c_void_p), C.cast(bb, C.c_void_p)))
        return h

    def release(self, device=None):
        """destroy the device replicas (all of them, or the one on `device`)"""
        if not self._handles:
            return
        from . import _lib
        for dev in [d for d in list(self._handles) if device is None or d == device]:
            with _lib.on_device(dev):
                _lib.lib.gpbt_emulator_destroy(self._handles.pop(dev))

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass   # interpreter shutdown: the library may already be gone

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_handles"] = {}
        d.pop("_handle", None)
        return d

    def __setstate__(self, d):
        d = dict(d)
        d.pop("_handle", None)     # pickles written before the per-device table existed
        d["_handles"] = {}
        self.__dict__.update(d)
