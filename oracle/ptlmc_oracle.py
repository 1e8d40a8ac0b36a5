"""
CPU oracle for the device-resident PTLMC iteration loop  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in NumPy, the loop of the reference's parallel-tempering sampler (src/mcmc.py:623-671, the branch
without gradients) and its temperature exchange (src/mcmc.py:679-693) with the random draws of the device kernels
(csrc/ptlmc.cuh): Philox4x32-10 keyed by (seed; iteration, index, purpose) instead of NumPy's global generator.

Pin: the loop's arithmetic is the reference's (`gpbt_b200/ptlmc.py` reproduces the unmodified reference's seeded
chains with the same statements, tests/test_ptlmc.py); `exchange_sweep` below is checked against that pinned
`temp_exchange_python` on identical slots and uniforms (tests/test_host_logic.py).  What is NOT the reference's is
the random stream -- by construction.

    normals   (k, i * ceil(p/2) + e2, 16): u1 = 1 - u(w0, w1), u2 = u(w2, w3); Box-Muller -> dims 2 e2, 2 e2 + 1
    accept    (k, i, 17): log u(w0, w1)
    sweep s   (k, j, 20 + s), j = 0 .. n - 1: slot rt = 1 + floor(u(w0, w1) (n - 1)), log u(w2, w3)
"""
import numpy as np

from .ensemble_oracle import _u01, philox4x32

TAG_NORMAL, TAG_ACCEPT, TAG_SWEEP, SWEEPS = 16, 17, 20, 5


def stride_of(tau):
    e = np.exp(2.0 * tau)
    return 2.0 * (1.0 + (e - 1.0) / (e + 1.0))


def normals(seed, k, n, p):
    p2 = (p + 1) // 2
    xi = np.zeros((n, 2 * p2))
    for i in range(n):
        for e2 in range(p2):
            r = philox4x32(seed, k, i * p2 + e2, TAG_NORMAL)
            u1, u2 = 1.0 - _u01(r[0], r[1]), _u01(r[2], r[3])
            rad = np.sqrt(-2.0 * np.log(u1))
            xi[i, 2 * e2] = rad * np.cos(6.283185307179586 * u2)
            xi[i, 2 * e2 + 1] = rad * np.sin(6.283185307179586 * u2)
    return xi[:, :p]


def sweep_draws(seed, k, sweep, n):
    """slots [n] in [1, n) and log-uniforms [n] of exchange sweep `sweep` of iteration k"""
    slots, logu = np.zeros(n, dtype=np.int64), np.zeros(n)
    for j in range(n):
        r = philox4x32(seed, k, j, TAG_SWEEP + sweep)
        slots[j] = min(1 + int(_u01(r[0], r[1]) * (n - 1)), n - 1)
        with np.errstate(divide="ignore"):
            logu[j] = np.log(_u01(r[2], r[3]))
    return slots, logu


def exchange_sweep(lp, temps, order, slots, logu):
    """src/mcmc.py:683-692 on given draws; order is updated in place"""
    for rt, lu in zip(slots, logu):
        gap = 1.0 / temps[rt - 1] - 1.0 / temps[rt]
        if (lp[order[rt]] - lp[order[rt - 1]]) * gap > lu:
            order[rt - 1], order[rt] = order[rt], order[rt - 1]
    return order


def run(logpost, theta0, temps, root, n_hot, n_tune, n_keep, seed, goal=0.25, tau=-1.0):
    """The device loop.  logpost(theta [n, p]) -> lp [n].  Returns dict(theta = recorded T = 1 chains
    [n - n_hot, n_keep, p], state, lp, tau, accepted, takes = accept pattern per iteration [iters, n])."""
    theta = np.array(theta0, dtype=np.float64, copy=True)
    temps = np.asarray(temps, dtype=np.float64).reshape(-1)
    n, p = theta.shape
    lp = np.asarray(logpost(theta), dtype=np.float64).copy()
    cbrt = temps ** (1.0 / 3.0)
    saved = np.zeros((n - n_hot, n_keep, p))
    takes = np.zeros((n_tune + n_keep, n), dtype=bool)
    hits, accepted, stride = 0.0, 0, stride_of(tau)
    for k in range(n_tune + n_keep):
        xi = normals(seed, k, n, p)
        scale = 1.4142135623730951 * (stride * cbrt)
        prop = theta + scale[:, None] * (xi @ root)
        lpn = np.asarray(logpost(prop), dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            logu = np.array([np.log(_u01(*philox4x32(seed, k, i, TAG_ACCEPT)[:2])) for i in range(n)])
            take = logu < (lpn - lp) / temps
        takes[k] = take
        theta[take] = prop[take]
        lp[take] = lpn[take]
        hits += take.sum() / n
        if k >= n_tune:
            accepted += int(take[n_hot:].sum())
        order = np.arange(n)
        if n > 1:
            for s in range(SWEEPS):
                exchange_sweep(lp, temps, order, *sweep_draws(seed, k, s, n))
        theta, lp = theta[order], lp[order]
        if k < n_tune and k % 10 == 0:
            tau = tau + 1.0 / np.sqrt(1.0 + k / 10.0) * (hits / 10.0 - goal)
            stride = stride_of(tau)
            hits = 0.0
        elif k >= n_tune:
            saved[:, k - n_tune, :] = theta[n_hot:]
    return {"theta": saved, "state": theta, "lp": lp, "tau": tau, "accepted": accepted, "takes": takes}
