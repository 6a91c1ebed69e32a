"""Prototype of the tree-structured ground-state solver (numpy), validated against numpy.linalg.eigh.
H is a tree: diagonal delta_i (relative to H_00) and one coupling beta_i to parent(i) < i.
Lowest eigenvalue: Laguerre iteration on det(H - x) = prod of the elimination pivots d_i(x), started left of the
spectrum (monotone, cubic, all pivots stay positive); eigenvector: inverse iteration with the same factorisation."""
import numpy as np

def make(S, seed, tiny=False, maxdepth=3):
    r = np.random.default_rng(seed)
    parent = np.full(S, -1); level = np.zeros(S, int)
    for i in range(1, S):
        while True:
            p = int(r.integers(0, i))
            if level[p] < maxdepth: break
        parent[i] = p; level[i] = level[p] + 1
    delta = np.concatenate([[0.0], r.uniform(-30, 120, S - 1)])
    beta = np.concatenate([[0.0], r.uniform(-60, -1, S - 1)])
    if tiny and S > 1:
        beta[r.integers(1, S)] = 1e-30
    return parent, level, delta, beta

def dense(parent, delta, beta):
    S = len(delta); H = np.diag(delta)
    for i in range(1, S): H[i, parent[i]] = H[parent[i], i] = beta[i]
    return H

def solve(parent, level, delta, beta, x0=None):
    S = len(delta)
    children = [[] for _ in range(S)]
    for i in range(1, S): children[parent[i]].append(i)
    maxlev = level.max()
    d = np.zeros(S); d1 = np.zeros(S); d2 = np.zeros(S); invd = np.zeros(S)
    def ev(x):
        for L in range(maxlev, -1, -1):
            for i in range(S):
                if level[i] != L: continue
                s0 = s1 = s2 = 0.0
                for c in children[i]:
                    b2 = beta[c] ** 2; r = invd[c]
                    s0 += b2 * r; s1 += b2 * d1[c] * r * r; s2 += b2 * (d2[c] * r * r - 2 * d1[c] ** 2 * r ** 3)
                d[i] = delta[i] - x - s0; d1[i] = -1.0 + s1; d2[i] = s2
                invd[i] = 1.0 / d[i] if d[i] != 0 else np.inf
        return (d > 0).all()
    absb = np.abs(beta)
    gers = np.array([delta[i] - absb[i] - sum(absb[c] for c in children[i]) for i in range(S)])
    lo = gers.min() - 1e-3 * (1 + abs(gers.min())); hi = np.inf
    tol = 1e-10 * max(np.abs(delta).max(), abs(lo), 1.0)     # the eigenvector iteration and the Rayleigh quotient finish the job
    if S == 1: return delta[0], np.ones(1), 0
    x = lo if (x0 is None or x0 <= lo) else x0
    n_eval = 0; retreat = 0
    for it in range(200):
        ok = ev(x); n_eval += 1
        if not ok:
            # right of (or, to rounding, on) the lowest eigenvalue: retreat geometrically from the invalid point --
            # a cubic step that lands on the root itself needs one retreat of tol/2, not a bisection of (lo, hi)
            hi = min(hi, x)
            if hi - lo <= tol: x = lo; ev(x); n_eval += 1; break
            x = max(hi - 0.5 * tol * 8.0 ** retreat, 0.5 * (lo + hi))
            retreat += 1
            continue
        lo = x
        if hi - lo <= tol: break
        G = (d1 * invd).sum(); S2 = (d2 * invd - (d1 * invd) ** 2).sum(); Hh = -S2
        n = float(S)
        disc = max((n - 1) * (n * Hh - G * G), 0.0)
        a = n / (G - np.sqrt(disc))           # G < 0 left of the spectrum: a < 0
        xn = x - a
        if xn - x <= tol: break
        if not (xn < hi): xn = 0.5 * (lo + hi)
        x = xn
    mu = x
    y = np.full(S, 1.0 / np.sqrt(S))
    for rep in range(4):
        b = y.copy()
        for L in range(maxlev, -1, -1):
            for i in range(S):
                if level[i] == L:
                    b[i] = b[i] - sum(beta[c] * (b[c] * invd[c]) for c in children[i])
        ynew = np.zeros(S)
        for L in range(0, maxlev + 1):
            for i in range(S):
                if level[i] == L:
                    ynew[i] = (b[i] - (beta[i] * ynew[parent[i]] if i > 0 else 0.0)) * invd[i]
        ynew /= np.linalg.norm(ynew)
        done = np.abs(np.abs(ynew) - np.abs(y)).max() < 1e-15
        y = ynew
        n_eval += 1
        if done: break
    mu = (delta * y * y).sum() + 2 * sum(beta[i] * y[i] * y[parent[i]] for i in range(1, S))
    return mu, y, n_eval

if __name__ == "__main__":
    worst = 0; worstv = 0; evs = []; evw = []
    for seed in range(600):
        S = int(np.random.default_rng(seed).integers(1, 81))
        p, l, dl, b = make(S, seed, tiny=(seed % 7 == 0))
        mu, y, n = solve(p, l, dl, b)
        w, v = np.linalg.eigh(dense(p, dl, b))
        v0 = v[:, 0] * np.sign(v[:, 0] @ y)
        gap = w[1] - w[0] if S > 1 else 1.0
        e1 = abs(mu - w[0]); e2 = np.abs(v0 - y).max() * min(gap, 1.0)
        if e1 > 1e-11 or e2 > 1e-11: print("BAD seed", seed, S, e1, e2, gap)
        worst = max(worst, e1); worstv = max(worstv, e2); evs.append(n)
        dl2 = dl + np.random.default_rng(seed + 1).uniform(-.3, .3, S) * (np.arange(S) > 0)
        mu2, y2, n2 = solve(p, l, dl2, b, x0=mu - 1.0)
        w2 = np.linalg.eigvalsh(dense(p, dl2, b))
        worst = max(worst, abs(mu2 - w2[0])); evw.append(n2)
    print("max |mu - eigh| = %.2e   max gap-weighted |c - eigh| = %.2e   evals cold: mean %.1f max %d   warm: mean %.1f max %d"
          % (worst, worstv, np.mean(evs), max(evs), np.mean(evw), max(evw)))

def trace(seed, warm=False):
    S = int(np.random.default_rng(seed).integers(1, 81))
    p, l, dl, b = make(S, seed, tiny=(seed % 7 == 0))
    import builtins
    w = np.linalg.eigvalsh(dense(p, dl, b))
    print("S", S, "lam0", w[0], "lam1", w[1] if S > 1 else None)
    return solve(p, l, dl, b)
