"""
Golden vectors for the surmise path (EmulatorBAND.predict, /root/reference/src/emulator_BAND.py:386-478).

The arithmetic of that path lives in surmise 0.2.1 (requirements.txt:1), which is neither installed in the
build image nor vendored in the reference tree, so the committed tests cannot pin kernel kind 2 to the
library yet ("parity unpinned").  This script is the missing step: run it ONCE in an
environment that has `surmise==0.2.1` next to the reference and commit the two files it writes --

    python tests/golden/make_golden_band.py          ->  tests/golden/band_pcgp.npz, band_pcsk.npz

tests/test_band_golden.py then stops xfail-ing and checks the oracle restatement (CPU) and the CUDA path
(GPU) against them: PC-space mean / variance, observable-space mean / covariance (which also settles
whether covx() carries `extravar`) and Chain.log_posterior.

What is stored per method (PCGP and PCSK), all from the UNMODIFIED reference class:
    fit information of the surmise emulator (`emu.emu._info`): theta, pct (or pcti), scale, offset, extravar,
    and per PC hypcov, hypind, nug, Vh, pw, sig2;
    X [N, p] (1 % of the rows outside the box), EmulatorBAND.predict(X, return_cov=True) -> mean, cov,
    Chain.log_posterior(X), Chain.log_likelihood(X, finite=True), the experiment (y_exp, cov_exp), the box.
"""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import WORK, _import_reference, _import_synthetic  # noqa: E402

SHAPE = dict(p=6, n=60, m=24, N=48)


def fit_arrays(info):
    """the pieces of surmise's fit dictionary that predict() reads, as plain arrays"""
    emul = info["emulist"]
    out = {"theta": np.asarray(info["theta"], dtype=np.float64),
           "pct": np.asarray(info["pct"] if "pct" in info else info["pcti"], dtype=np.float64),
           "pct_key": np.array("pct" if "pct" in info else "pcti"),
           "scale": np.asarray(info["scale"], dtype=np.float64).reshape(-1),
           "offset": np.asarray(info["offset"], dtype=np.float64).reshape(-1),
           "extravar": np.asarray(info["extravar"], dtype=np.float64).reshape(-1),
           "n_pc": np.array(len(emul))}
    for k, e in enumerate(emul):
        out["pc%d_hypcov" % k] = np.asarray(e["hypcov"], dtype=np.float64).reshape(-1)
        out["pc%d_hypind" % k] = np.array(int(e.get("hypind", k)))
        out["pc%d_nug" % k] = np.array(float(np.squeeze(e["nug"])))
        out["pc%d_Vh" % k] = np.asarray(e["Vh"], dtype=np.float64)
        out["pc%d_pw" % k] = np.asarray(e["pw"], dtype=np.float64).reshape(-1)
        out["pc%d_sig2" % k] = np.array(float(np.squeeze(e["sig2"])))
    return out


def main():
    try:
        import surmise  # noqa: F401
    except ImportError:
        raise SystemExit("surmise is not importable here: install surmise==0.2.1 (requirements.txt of the "
                         "reference) and run this script again; nothing was written")
    _, Chain, _ = _import_reference()
    from src.emulator_BAND import EmulatorBAND
    syn = _import_synthetic()
    p, n, m, N = SHAPE["p"], SHAPE["n"], SHAPE["m"], SHAPE["N"]
    paths = syn.write_fixture(os.path.join(WORK, "band"), p, n, m)
    X = syn.walkers(p, N, seed=21)
    for method in ("PCGP", "PCSK"):
        emu = EmulatorBAND(training_set_path=paths["train"], parameter_file=paths["par"], method=method)
        emu.trainEmulatorAutoMask()
        lo, hi = syn.box(p)
        inside = np.all((X > lo) & (X < hi), axis=1)
        mean, cov = emu.predict(X[inside], return_cov=True)
        epath = os.path.join(WORK, "band", "emu_%s.pkl" % method)
        import dill
        with open(epath, "wb") as fh:
            dill.dump(emu, fh)
        ch = Chain(mcmc_path=os.path.join(WORK, "mcmc", "band.pkl"), expdata_path=paths["exp"],
                   model_parafile=paths["par"])
        ch.loadEmulator([epath])
        out = fit_arrays(emu.emu._info)
        out.update(method=np.array(method), surmise_version=np.array(getattr(surmise, "__version__", "unknown")),
                   X=X, inside=inside, mean=mean, cov=cov, lo=lo, hi=hi, y_exp=ch.expdata, cov_exp=ch.expdata_cov,
                   lp_posterior=ch.log_posterior(X), lp_like_finite=ch.log_likelihood(X, finite=True))
        dst = os.path.join(HERE, "band_%s.npz" % method.lower())
        np.savez_compressed(dst, **out)
        print("wrote", dst)


if __name__ == "__main__":
    main()
