"""Executable statement of the ECDF index of csrc/plugin.cuh ("ECDF"): levels sampled with stride 9, 64-byte nodes in probe order
[e3 e6 | e1 e2 | e4 e5 | e7 e8], a power-of-two top level searched by integer bisection on split 32-bit words, two 16-byte probes per
level, and the running bracket (largest value < x, smallest value >= x) that must end as (K[j], K[j+1]) of Interpolations.jl's
`searchsortedfirst - 1` rule.  This numpy model follows the device code line by line and is compared with np.searchsorted on
continuous, duplicate-heavy and tiny tables; the device code itself is pinned against the oracle by the GPU tests."""
import numpy as np
import pytest

STRIDE, NODE = 9, 8
PERM = [3, 6, 1, 2, 4, 5, 7, 8]


def build(K, top_max):
    levels, cur = [], K
    while cur.size > top_max:
        n_nodes = (cur.size + STRIDE - 1) // STRIDE
        nodes = np.full((n_nodes, NODE), np.inf)
        for s, w in enumerate(PERM):
            idx = np.arange(n_nodes) * STRIDE + w
            ok = idx < cur.size
            nodes[ok, s] = cur[idx[ok]]
        levels.append(nodes)
        cur = cur[::STRIDE].copy()
    P = 2
    while P < cur.size:
        P *= 2
    top = np.full(P, np.inf)
    top[:cur.size] = cur
    bits = top.view(np.uint64)
    return levels, (bits >> np.uint64(32)).astype(np.uint32), (bits & np.uint64(0xffffffff)).astype(np.uint32), top


def lookup(K, levels, th, tl, top, rho):
    kmax = K[-1]
    x = kmax if rho > kmax else (0.0 if rho < 0.0 else rho)
    x = x + 0.0
    xb = np.float64(x).view(np.uint64)
    xh, xl = np.uint32(xb >> np.uint64(32)), np.uint32(xb & np.uint64(0xffffffff))
    P = th.size
    c, step = 0, P >> 1
    while step >= 1:
        i = c + step - 1
        lt = th[i] < xh or (th[i] == xh and tl[i] < xl)
        c += step if lt else 0
        step >>= 1
    lt = th[c] < xh or (th[c] == xh and tl[c] < xl)
    c += 1 if lt else 0
    if c == 0:
        return 0, None, None
    lo, hi = top[c - 1], (top[c] if c < P else np.inf)
    lb = c
    for nodes in reversed(levels):
        nd = nodes[lb - 1]
        s0, s1 = nd[0], nd[1]
        c1 = int(s0 < x) + int(s1 < x)
        p0, p1 = nd[2 + 2 * c1], nd[3 + 2 * c1]
        c2 = int(p0 < x) + int(p1 < x)
        lo_s = lo if c1 == 0 else (s0 if c1 == 1 else s1)
        hi_s = s0 if c1 == 0 else (s1 if c1 == 1 else hi)
        lo = lo_s if c2 == 0 else (p0 if c2 == 1 else p1)
        hi = p0 if c2 == 0 else (p1 if c2 == 1 else hi_s)
        lb = (lb - 1) * STRIDE + 1 + 3 * c1 + c2
    return lb, lo, hi


@pytest.mark.parametrize("kind,n,top_max", [("continuous", 20_000, 64), ("continuous", 5_000, 2048), ("duplicates", 30_000, 128), ("integers", 9_000, 64),
                                            ("tiny", 3, 64), ("continuous", 81 * 64 + 1, 64), ("continuous", 729, 8)])
def test_index_lookup_equals_searchsorted(kind, n, top_max):
    rng = np.random.default_rng(hash((kind, n)) % 2**32)
    if kind == "continuous":
        d = rng.gamma(2.0, 1.0, n)
    elif kind == "duplicates":
        d = np.round(rng.gamma(2.0, 1.0, n), 1) + 0.1
    elif kind == "integers":
        d = rng.integers(1, 40, n).astype(float) ** 2
    else:
        d = np.array([1.0, 2.0, 2.0])[:n]
    K = np.concatenate([[0.0], np.sort(d[d > 0]), [d.max() * 1.5]])
    levels, th, tl, top = build(K, top_max)
    L = K.size
    queries = np.concatenate([rng.choice(K, 400), rng.uniform(0, K[-1] * 1.2, 400), np.nextafter(rng.choice(K, 200), np.inf),
                              np.nextafter(rng.choice(K[1:], 200), -np.inf), [0.0, -0.0, -1.0, K[-1], K[-1] * 2, K[1], K[-2]]])
    for rho in queries:
        lb, lo, hi = lookup(K, levels, th, tl, top, float(rho))
        x = min(max(float(rho), 0.0), K[-1])
        want_lb = int(np.searchsorted(K, x, side="left"))                  # number of knots < x  (searchsortedfirst - 1, 0-based)
        assert lb == want_lb, (rho, lb, want_lb)
        if lb > 0:
            j = lb - 1
            assert j <= L - 2 and lo == K[j] and hi == K[j + 1], (rho, j, lo, hi, K[j], K[j + 1])
