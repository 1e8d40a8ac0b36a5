"""
CPU oracle for the device-resident ensemble sampler  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in NumPy, one step of the affine-invariant ensemble sampler the reference drives through
emcee (src/mcmc.py:68-92 LoggingEnsembleSampler, :372-412 Chain.run_mcmc; default move of
emcee.EnsembleSampler = StretchMove(a=2) on a random red/blue split).

Pin: emcee (requirements.txt: emcee>=3.1.4) is neither installed here nor vendored in the reference
tree, so this file follows the published algorithm -- Goodman & Weare (2010), as implemented by
emcee 3's RedBlueMove.propose / StretchMove.get_proposal:

    split the walkers at random into two sets; for each set S in turn, with C the other set:
        zz      = ((a - 1) u + 1)^2 / a                 u ~ U[0,1)    one per walker of S
        factors = (ndim - 1) log zz
        rint    ~ randint(len(C))                                      one per walker of S
        q       = C[rint] - (C[rint] - S) zz
        accept  where  factors + log_prob(q) - log_prob(S) > log(U[0,1))

"parity unpinned" against emcee itself (no golden vector can be generated without the package);
what IS pinned is the arithmetic of a step given the random draws, and the Philox4x32-10 generator
against the known-answer vectors of its publication (tests/test_host_logic.py).

The random draws are explicit inputs, so the device kernels can be compared draw for draw:
    u       [steps, 2, n_half, 2]   (stretch draw, accept draw) per half step and active walker
    partner [steps, 2, n_half]      index into the complementary set
    perm    [steps, n_walkers]      split: perm[:n_half] is the first set, perm[n_half:] the second
with n_half = ceil(n_walkers / 2).
"""
import numpy as np

M32 = 0xFFFFFFFF


def philox4x32(seed, ctr_lo, idx, tag, rounds=10):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11): counter (ctr_lo[31:0], ctr_lo[63:32], idx, tag),
    key (seed[31:0], seed[63:32]); returns four 32-bit words."""
    c = [ctr_lo & M32, (ctr_lo >> 32) & M32, idx & M32, tag & M32]
    k0, k1 = seed & M32, (seed >> 32) & M32
    for _ in range(rounds):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k0, p1 & M32, (p0 >> 32) ^ c[3] ^ k1, p0 & M32]
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c


def _u01(hi, lo):
    return float(((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0)


RANK_MAX_WALKERS = 2048   # above this the device splits with a keyed bijection instead of ranking keys


def feistel_perm(seed, step, nw):
    """perm[i] = 4-round Feistel network on the next power of four >= nw (round function: Philox word 0,
    tags 8..11), cycle-walked into [0, nw) -- csrc/ensemble.cuh split_feistel."""
    hbits = 1
    while (1 << (2 * hbits)) < nw:
        hbits += 1
    mask = (1 << hbits) - 1
    out = np.empty(nw, dtype=np.int32)
    for i in range(nw):
        x = i
        while True:
            l, r = x >> hbits, x & mask
            for rd in range(4):
                l, r = r, l ^ (philox4x32(seed, step, r, 8 + rd)[0] & mask)
            x = (l << hbits) | r
            if x < nw:
                break
        out[i] = x
    return out


def philox_streams(seed, first_step, nsteps, nw):
    """The draws the device generates for steps [first_step, first_step + nsteps): tag 0/1 = uniforms
    of the two half steps, 3/4 = partner draws, 2 = split keys (csrc/ensemble.cuh)."""
    n0 = (nw + 1) // 2
    u = np.zeros((nsteps, 2, n0, 2))
    partner = np.zeros((nsteps, 2, n0), dtype=np.int32)
    perm = np.zeros((nsteps, nw), dtype=np.int32)
    for s in range(nsteps):
        step = first_step + s
        if nw > RANK_MAX_WALKERS:
            perm[s] = feistel_perm(seed, step, nw)
        else:
            keys = []
            for j in range(nw):
                r = philox4x32(seed, step, j, 2)
                keys.append(((((r[0] << 32) | r[1]) & ~0xFFFFFF) | j, j))   # 40 random bits above the index
            perm[s] = [j for _, j in sorted(keys)]
        for half in range(2):
            ns = n0 if half == 0 else nw - n0
            nc = nw - ns
            for i in range(ns):
                w = philox4x32(seed, step, i, half)
                v = philox4x32(seed, step, i, 3 + half)
                u[s, half, i, 0] = _u01(w[0], w[1])
                u[s, half, i, 1] = _u01(w[2], w[3])
                partner[s, half, i] = ((((v[0] << 32) | v[1]) * nc) >> 64)
    return u, partner, perm


def fixed_split(nw):
    """perm of the non-randomised split: even walkers, then odd walkers (emcee's arange % 2)."""
    return np.concatenate((np.arange(0, nw, 2), np.arange(1, nw, 2))).astype(np.int32)


def stretch_run(log_prob, x0, lp0, u, partner, perm, a=2.0):
    """Runs len(u) steps from (x0, lp0).  Returns chain [steps, nw, p], lp [steps, nw], accepted [nw]."""
    x = np.array(x0, dtype=np.float64)
    lp = np.array(lp0, dtype=np.float64)
    nw, p = x.shape
    n0 = (nw + 1) // 2
    steps = len(u)
    chain, lps, accepted = np.empty((steps, nw, p)), np.empty((steps, nw)), np.zeros(nw, dtype=np.int64)
    for s in range(steps):
        for half in range(2):
            S = perm[s, :n0] if half == 0 else perm[s, n0:]
            Cc = perm[s, n0:] if half == 0 else perm[s, :n0]
            ns = len(S)
            if ns == 0:
                continue
            zz = ((a - 1.0) * u[s, half, :ns, 0] + 1.0) ** 2 / a
            factors = (p - 1.0) * np.log(zz)
            c = x[Cc[partner[s, half, :ns]]]
            q = c - (c - x[S]) * zz[:, None]
            new_lp = np.asarray(log_prob(q), dtype=np.float64)
            with np.errstate(invalid="ignore", divide="ignore"):
                acc = factors + new_lp - lp[S] > np.log(u[s, half, :ns, 1])
            x[S[acc]] = q[acc]
            lp[S[acc]] = new_lp[acc]
            accepted[S[acc]] += 1
        chain[s], lps[s] = x, lp
    return chain, lps, accepted
