#!/usr/bin/env python
"""How far outside a triangle's exact bounding box can a ray pass and still be accepted by the fp32 Moller-Trumbore rule?
(The production BVH culls with per-triangle boxes grown by `grow`; a hit is lost only if this distance exceeds `grow`.)
Random triangles of edge ~L at distance ~D from the ray origin, rays aimed at points within +-5e-3*L of the triangle's outline;
for every ray the rule accepts, the smallest box growth g that makes the reference's slab test (fp32, select form) pass is found by
bisection.  Prints the maximum g over all accepted rays, in units of D * 2^-24 (one ulp of the origin's coordinates is D * 2^-23)."""
import sys
import numpy as np
f32 = np.float32


def cross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - b[:, 1] * a[:, 2], a[:, 2] * b[:, 0] - b[:, 2] * a[:, 0], a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]], 1).astype(f32)


def dot(a, b):
    p = (a * b).astype(f32)
    return ((p[:, 0] + p[:, 1]).astype(f32) + p[:, 2]).astype(f32)


def mt(o, d, v0, v1, v2):
    e1, e2 = (v1 - v0).astype(f32), (v2 - v0).astype(f32)
    p = cross(d, e2)
    det = dot(e1, p)
    ok = np.abs(det) >= f32(1e-8)
    inv = (f32(1.0) / np.where(ok, det, f32(1))).astype(f32)
    s = (o - v0).astype(f32)
    u = (dot(s, p) * inv).astype(f32)
    ok &= (u >= 0) & (u <= 1)
    q = cross(s, e1)
    v = (dot(d, q) * inv).astype(f32)
    ok &= (v >= 0) & ((u + v).astype(f32) <= 1)
    t = (dot(e2, q) * inv).astype(f32)
    return ok & (t > f32(1e-4))


def slab(o, d, lo, hi):
    inv = (f32(1.0) / d).astype(f32)
    neg = inv < 0
    t0 = ((np.where(neg, hi, lo) - o).astype(f32) * inv).astype(f32)
    t1 = ((np.where(neg, lo, hi) - o).astype(f32) * inv).astype(f32)
    tmin = np.maximum(f32(0), t0.max(1))
    tmax = np.minimum(f32(3.4e38), t1.min(1))
    return ~(tmax < tmin)


def run(D, L, n, rng, aligned=False):
    c = rng.normal(0, 1, (n, 3)); c = (c / np.linalg.norm(c, axis=1, keepdims=True) * D * rng.uniform(0.5, 1.0, (n, 1)))
    if aligned:
        # Marching-Cubes-like: right triangles in a grid plane, legs along two axes, vertices on a lattice of pitch L / 2
        c = np.round(c / (L / 2)) * (L / 2)
        ax = rng.integers(0, 3, n)                        # the plane's normal axis
        a1, a2 = (ax + 1) % 3, (ax + 2) % 3
        s1, s2 = rng.choice([-1.0, 1.0], n), rng.choice([-1.0, 1.0], n)
        v0 = c.copy(); v1 = c.copy(); v2 = c.copy()
        v1[np.arange(n), a1] += s1 * L; v2[np.arange(n), a2] += s2 * L
        v = [v0.astype(f32), v1.astype(f32), v2.astype(f32)]
    else:
        v = [(c + rng.normal(0, L, (n, 3))).astype(f32) for _ in range(3)]
    o = (rng.normal(0, 0.1 * D, (n, 3))).astype(f32)
    # aim at a point on the outline, pushed outwards / inwards by a few 1e-3 L
    k = rng.integers(0, 3, n); a = rng.uniform(0, 1, (n, 1))
    va = np.choose(k[:, None], [v[0], v[1], v[2]]); vb = np.choose(k[:, None], [v[1], v[2], v[0]])
    target = va * (1 - a) + vb * a + rng.normal(0, 5e-3 * L, (n, 3))
    d = (target - o); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(f32)
    acc = mt(o, d, *v)
    lo = np.minimum(np.minimum(v[0], v[1]), v[2]); hi = np.maximum(np.maximum(v[0], v[1]), v[2])
    o, d, lo, hi = o[acc], d[acc], lo[acc], hi[acc]
    need = np.zeros(len(o))
    miss = ~slab(o, d, lo, hi)
    g_lo, g_hi = np.zeros(len(o)), np.full(len(o), 1e-2 * L)
    for _ in range(40):
        g = 0.5 * (g_lo + g_hi)
        ok = slab(o, d, (lo - g[:, None].astype(f32)).astype(f32), (hi + g[:, None].astype(f32)).astype(f32))
        g_hi = np.where(ok, g, g_hi); g_lo = np.where(ok, g_lo, g)
    need = np.where(miss, g_hi, 0.0)
    return int(acc.sum()), int(miss.sum()), float(need.max()) if len(need) else 0.0


if __name__ == "__main__":
    rng = np.random.default_rng(5)
    for aligned in (False, True):
      print("axis-aligned lattice triangles (Marching-Cubes-like)" if aligned else "random triangles")
      for D, L in ((1.2, 0.0078), (2550.0, 10.0), (8800.0, 10.0), (35000.0, 10.0), (460.0, 1.0), (1840.0, 1.0)):
        tot_acc = tot_miss = 0; worst = 0.0
        for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
            a, m, w = run(D, L, 400000, rng, aligned)
            tot_acc += a; tot_miss += m; worst = max(worst, w)
        print("D %-8g L %-7g accepted %8d  of which outside the exact box %6d  max growth needed %.3e = %.1f x D*2^-24   (grow at ext=D/1.2: %.3e)" %
              (D, L, tot_acc, tot_miss, worst, worst / (D * 2.0 ** -24), D / 1.2 * 2.0 ** -18))
