"""
Generate golden vectors by running the UNMODIFIED reference (/root/reference/src) on synthetic
fixtures.  Runs only in the build container (the GPU box has no /root/reference); the .npz files
it writes are committed and are what tests/ compare against.

    python tests/golden/make_golden.py [--only c1_rbf,...] [--c2]

Reference entry points exercised (file:line under /root/reference):
    Emulator.__init__/trainEmulator/predict        src/emulator.py:50-99, 257-363, 465-605
    Chain.__init__/loadEmulator/_predict            src/mcmc.py:104-166
    Chain.log_posterior / log_likelihood            src/mcmc.py:188-222, 261-299
    mvn_loglike                                     src/mcmc.py:23-65
emcee / pocomc are not installed here; src/mcmc.py imports them at module top, so empty stub
modules are registered first (none of the exercised functions touch them).
"""
import argparse
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
WORK = "/tmp/gpbt_golden_work"


def _import_reference():
    os.environ["WORKDIR"] = WORK
    os.environ.setdefault("LOGLEVEL", "warning")
    os.makedirs(os.path.join(WORK, "mcmc"), exist_ok=True)
    sys.path.insert(0, "/root/reference")
    em = types.ModuleType("emcee")
    em.EnsembleSampler = type("EnsembleSampler", (), {"__init__": lambda s, *a, **k: None})
    pm = types.ModuleType("pocomc")
    pm.Prior = object
    pm.Sampler = object
    sys.modules.setdefault("emcee", em)
    sys.modules.setdefault("pocomc", pm)
    from src.emulator import Emulator
    from src.mcmc import Chain, mvn_loglike
    return Emulator, Chain, mvn_loglike


def _import_synthetic():
    spec = importlib.util.spec_from_file_location(
        "gpbt_synthetic", os.path.join(ROOT, "gpbayestools-hic_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def extract_state(emu, kind, prefix=""):
    """Trained quantities of a reference Emulator as plain arrays (see oracle/gp_oracle.py)."""
    gps = emu.gps
    n = gps[0].X_train_.shape[0]
    tril = np.tril_indices(n)
    st = {
        "kind": np.array(kind),
        "Xtr": np.array(gps[0].X_train_, dtype=np.float64),
        "ell": np.array([np.broadcast_to(g.kernel_.k1.k2.length_scale,
                                         gps[0].X_train_.shape[1]) for g in gps], dtype=np.float64),
        "c": np.array([g.kernel_.k1.k1.constant_value for g in gps], dtype=np.float64),
        "sn": np.array([g.kernel_.k2.noise_level for g in gps], dtype=np.float64),
        "alpha": np.array([g.alpha_ for g in gps], dtype=np.float64),
        "Lpacked": np.array([g.L_[tril] for g in gps], dtype=np.float64),
        "no_pca": np.array(bool(emu.perform_no_PCA_)),
        "exp_diag": np.array(bool(emu.exp_and_cov_diagonal_)),
        "mu": np.array(emu.scaler.mean_, dtype=np.float64),
        "scale": np.array(emu.scaler.scale_, dtype=np.float64),
    }
    if not emu.perform_no_PCA_:
        st["A"] = np.array(emu._trans_matrix[:emu.npc], dtype=np.float64)
        st["Ctrunc"] = np.array(emu._cov_trunc, dtype=np.float64)
    return {prefix + k: v for k, v in st.items()}


def trafo_arrays(emu, prefix):
    """scalers / PCAs of the reference's parameterTrafoPCA (src/emulator.py:79-99, 129-241)"""
    out = {prefix + "trafo_p_in": np.array(emu.design_points_org_.shape[1])}
    for tag, idx in (("bulk", emu.indices_zeta_s_parameters), ("shear", emu.indices_eta_s_parameters),
                     ("yloss", emu.indices_yloss_parameters)):
        sc, pc = getattr(emu, "paramTrafoScaler_" + tag), getattr(emu, "paramTrafoPCA_" + tag)
        out.update({prefix + "trafo_%s_idx" % tag: np.array(idx), prefix + "trafo_%s_smean" % tag: sc.mean_,
                    prefix + "trafo_%s_sscale" % tag: sc.scale_, prefix + "trafo_%s_pmean" % tag: pc.mean_,
                    prefix + "trafo_%s_comp" % tag: pc.components_})
    return out


def make_case(name, syn, Emulator, Chain, mvn_loglike, p, n, emus, N, seed, keep_cov_rows,
              store_L=True, extra_std_scale=0.05, shift=0.0):
    """emus: list of dicts(m, q, kind, logTrafo, exp_diag, no_pca). One Chain over all of them."""
    import dill
    import pickle
    wd = os.path.join(WORK, name)
    os.makedirs(wd, exist_ok=True)
    m_total = sum(e["m"] for e in emus)
    # one simulator with m_total observables; each emulator trains on its slice of them
    full = syn.training_dict(p, n, m_total)
    exp = syn.experiment_dict(p, m_total)
    par_text = syn.parameter_file_text(p)
    if shift:
        # keep every coordinate away from 0 (the zeta/s widths sigma_+- of parameterTrafoPCA divide):
        # the whole box moves by `shift`, the simulator still sees the unshifted coordinates
        for v in full.values():
            v["parameter"] = v["parameter"] + shift
        lo_, hi_ = syn.box(p)
        par_text = "".join("par%d: p%d, %r, %r\n" % (d, d, float(lo_[d] + shift), float(hi_[d] + shift))
                           for d in range(p))
    par_path = os.path.join(wd, "par.txt")
    with open(par_path, "w") as f:
        f.write(par_text)
    exp_path = os.path.join(wd, "exp.pkl")
    with open(exp_path, "wb") as f:
        pickle.dump(exp, f)

    out = {"n_emu": np.array(len(emus))}
    emu_paths, ref_emus, o = [], [], 0
    for e_i, e in enumerate(emus):
        sl = slice(o, o + e["m"])
        o += e["m"]
        sub = {k: {"parameter": v["parameter"], "obs": v["obs"][:, sl]} for k, v in full.items()}
        tp = os.path.join(wd, "train%d.pkl" % e_i)
        with open(tp, "wb") as f:
            pickle.dump(sub, f)
        emu = Emulator(training_set_path=tp, parameter_file=par_path, npc=e["q"],
                       logTrafo=e.get("logTrafo", False),
                       exp_and_cov_diagonal=e.get("exp_diag", False),
                       perform_no_PCA=e.get("no_pca", False),
                       parameterTrafoPCA=e.get("param_trafo", False))
        emu.trainEmulator([True] * emu.nev, kernel_type=e["kind"])
        ep = os.path.join(wd, "emu%d.pkl" % e_i)
        with open(ep, "wb") as f:
            dill.dump(emu, f)
        emu_paths.append(ep)
        ref_emus.append(emu)
        st = extract_state(emu, e["kind"], prefix="e%d_" % e_i)
        if not store_L:
            st.pop("e%d_Lpacked" % e_i)
        out.update(st)
        if e.get("param_trafo", False):
            out.update(trafo_arrays(emu, "e%d_" % e_i))

    X = syn.walkers(p, N, seed=seed) + shift
    lo, hi = syn.box(p)
    lo, hi = lo + shift, hi + shift
    inside = np.all((X > lo) & (X < hi), axis=1)
    Xin = X[inside]
    out.update(X=X, lo=lo, hi=hi, inside=inside)

    # boundary #1: Emulator.predict(X, return_cov=True, extra_std=arr) per emulator, with a
    # NON-zero extra_std so that path is pinned too
    extra = extra_std_scale * Xin[:, -1]
    out["extra_std"] = extra
    for e_i, emu in enumerate(ref_emus):
        # per-GP sklearn outputs (PC space): mean and diag of the predictive covariance
        # (not for parameterTrafoPCA emulators: their GPs live in the transformed parameter space and
        # the reference has no stand-alone transform to feed them with)
        if getattr(emu, "parameterTrafoPCA_", False):
            rows = Xin[:max(keep_cov_rows, 1)]
            mean, cov = emu.predict(rows, return_cov=True, extra_std=extra[:len(rows)])
            out["e%d_mean_x" % e_i] = mean
            out["e%d_cov_x" % e_i] = cov
            out["e%d_mean0" % e_i] = emu.predict(Xin[:256], return_cov=False)
            continue
        zm = np.stack([g.predict(Xin, return_cov=False) for g in emu.gps], axis=1)
        if len(Xin) <= 512:
            zv = np.stack([g.predict(Xin, return_cov=True)[1].diagonal() for g in emu.gps], axis=1)
        else:
            zv = np.concatenate([np.stack([g.predict(Xin[s:s + 256], return_cov=True)[1].diagonal()
                                           for g in emu.gps], axis=1)
                                 for s in range(0, len(Xin), 256)], axis=0)
        out["e%d_z_mean" % e_i] = zm
        out["e%d_z_var" % e_i] = zv
        rows = Xin[:max(keep_cov_rows, 1)]
        mean, cov = emu.predict(rows, return_cov=True, extra_std=extra[:len(rows)])
        out["e%d_mean_x" % e_i] = mean
        out["e%d_cov_x" % e_i] = cov
        mean0 = emu.predict(Xin[:256], return_cov=False)
        out["e%d_mean0" % e_i] = mean0

    # boundary #2: Chain
    ch = Chain(mcmc_path=os.path.join(WORK, "mcmc", "chain.pkl"), expdata_path=exp_path,
               model_parafile=par_path)
    ch.loadEmulator(emu_paths)
    out["y_exp"] = np.array(ch.expdata, dtype=np.float64)
    out["cov_exp"] = np.array(ch.expdata_cov, dtype=np.float64)
    pm, pc = ch._predict(Xin[:keep_cov_rows], 0.0)
    out["chain_mean"] = pm
    out["chain_cov"] = pc
    out["lp_posterior"] = ch.log_posterior(X)
    out["lp_like_finite"] = ch.log_likelihood(X, finite=True)
    out["lp_like"] = ch.log_likelihood(X)
    # mvn_loglike pin on the first rows' (dY, cov)
    dY = pm - ch.expdata
    cv = pc + ch.expdata_cov
    # (mvn inputs are chain_mean - y_exp and chain_cov + cov_exp; not stored twice)
    out["mvn_val"] = np.array([mvn_loglike(a, b) for a, b in zip(dY, cv)])
    # full-covariance variant (BASELINE config 4): add a PSD systematic term to expdata_cov
    ch.expdata_cov = ch.expdata_cov + syn.systematic_cov(m_total)
    # (cov_exp_sys = cov_exp + synthetic.systematic_cov(m_total); deterministic, not stored)
    out["lp_posterior_sys"] = ch.log_posterior(X)

    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-12s -> %s  (%.1f KB)  lp[:3]=%s" % (name, path, os.path.getsize(path) / 1024,
                                                  out["lp_posterior"][:3]))


CASES = {
    # name: (p, n, emus, N, seed, keep_cov_rows)
    "c1_rbf": (5, 100, [dict(m=50, q=10, kind="RBF")], 128, 3, 16),
    "c1_matern": (5, 100, [dict(m=50, q=10, kind="Matern")], 64, 4, 8),
    "c1_logexp": (5, 100, [dict(m=50, q=10, kind="RBF", logTrafo=True, exp_diag=True)], 64, 5, 8),
    "c1_nopca": (5, 100, [dict(m=12, q=12, kind="RBF", no_pca=True)], 64, 6, 8),
    "c1_multi": (5, 100, [dict(m=30, q=8, kind="RBF"), dict(m=20, q=6, kind="Matern")], 64, 7, 8),
    "odd_shape": (3, 37, [dict(m=13, q=5, kind="RBF")], 33, 8, 8),
    # 20 parameters with the reference's parameterTrafoPCA groups (columns 2-4, 12-14, 15-18)
    "p20_trafo": (20, 90, [dict(m=16, q=6, kind="RBF", param_trafo=True)], 64, 9, 8),
}
SHIFT = {"p20_trafo": 0.05}
C2_CASE = {"c2_rbf": (17, 500, [dict(m=300, q=20, kind="RBF")], 1024, 3, 1),
           # the config-2 shape trained with the reference's kernel_type="Matern" (kernel (a) kind 1 at n = 500)
           "c2_matern": (17, 500, [dict(m=300, q=20, kind="Matern")], 256, 4, 1)}
STORE_L = set()   # cases above n = 100 whose file should carry the reference's L_ (1 MB per PC at n = 500)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--c2", action="store_true", help="also the config-2 shape (minutes)")
    a = ap.parse_args()
    Emulator, Chain, mvn_loglike = _import_reference()
    syn = _import_synthetic()
    cases = dict(CASES)
    if a.c2:
        cases.update(C2_CASE)
    only = [s for s in a.only.split(",") if s]
    for name, (p, n, emus, N, seed, keep) in cases.items():
        if only and name not in only:
            continue
        make_case(name, syn, Emulator, Chain, mvn_loglike, p, n, emus, N, seed, keep,
                  store_L=(n <= 100 or name in STORE_L), shift=SHIFT.get(name, 0.0))


if __name__ == "__main__":
    main()
