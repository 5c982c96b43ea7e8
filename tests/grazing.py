"""Adversarial rays for the tight BVH: almost in the plane of a triangle, aimed at (and just past) its
edges from far away, with the fp32 determinant spread log-uniformly over [0.9e-5, 1e-3] - the regime where the
reference's Moeller-Trumbore test (geometry.cu:38-65) accepts hits whose real intersection point lies outside the
triangle.  Plus a numpy restatement of that test in the reference's operation order (float32, round to nearest, no FMA)."""
import numpy as np

f32 = np.float32


def _cross(a, b):
    return np.stack([(a[:, 1] * b[:, 2]).astype(f32) - (a[:, 2] * b[:, 1]).astype(f32),
                     (a[:, 2] * b[:, 0]).astype(f32) - (a[:, 0] * b[:, 2]).astype(f32),
                     (a[:, 0] * b[:, 1]).astype(f32) - (a[:, 1] * b[:, 0]).astype(f32)], 1).astype(f32)


def _dot(a, b):
    return (((a[:, 0] * b[:, 0]).astype(f32) + (a[:, 1] * b[:, 1]).astype(f32)).astype(f32) + (a[:, 2] * b[:, 2]).astype(f32)).astype(f32)


def moeller_trumbore(o, d, v1, e1, e2):
    """Triangle::intersect (geometry.cu:38-65) on arrays: (accepted, t, u, v, det)."""
    P = _cross(d, e2)
    det = _dot(e1, P)
    ok = ~(np.abs(det) < f32(1e-5))
    with np.errstate(all="ignore"):
        inv = (f32(1) / det).astype(f32)
        T = (o - v1).astype(f32)
        u = (_dot(T, P) * inv).astype(f32)
        ok &= ~((u < 0) | (u > 1))
        Q = _cross(T, e1)
        v = (_dot(d, Q) * inv).astype(f32)
        ok &= ~((v < 0) | ((u + v).astype(f32) > 1))
        t = (_dot(e2, Q) * inv).astype(f32)
        ok &= t > f32(1e-5)
    return ok, t, u, v, det


def region_of(fs):
    """The origin region crt_create derives (crt_api.cu upload_scene): vertices + eye, grown by 5 % of the diagonal."""
    t = fs.tris
    v1 = t["v1"].astype(np.float64)
    pts = np.concatenate([v1, v1 + t["e1"], v1 + t["e2"], np.asarray(fs.camera["eye"], np.float64).reshape(1, 3)])
    lo, hi = pts.min(0).astype(f32), pts.max(0).astype(f32)
    m = 0.05 * np.linalg.norm(hi.astype(np.float64) - lo) + abs(float(np.asarray(fs.camera["apertureWidth"]).reshape(-1)[0])) + 1e-3
    return (lo.astype(np.float64) - m).astype(f32), (hi.astype(np.float64) + m).astype(f32)


def grazing_rays(fs, n, seed=5):
    """(o f32[n,3], d f32[n,3], k) - ray i grazes triangle k[i]; origins stay inside region_of(fs)."""
    rng = np.random.default_rng(seed)
    t = fs.tris
    lo, hi = region_of(fs)
    k = rng.integers(0, len(t), n)
    v1, e1, e2 = (t[f][k].astype(np.float64) for f in ("v1", "e1", "e2"))
    nrm = np.cross(e1, e2)
    A2 = np.linalg.norm(nrm, axis=1)
    nrm /= np.maximum(A2, 1e-300)[:, None]
    uu, vv = rng.uniform(-0.3, 1.3, n), rng.uniform(-0.3, 1.3, n)
    edge = rng.integers(0, 4, n)
    uu = np.where(edge == 0, rng.normal(0, 0.05, n), uu)
    vv = np.where(edge == 1, rng.normal(0, 0.05, n), vv)
    vv = np.where(edge == 2, 1 - uu + rng.normal(0, 0.05, n), vv)
    P = v1 + uu[:, None] * e1 + vv[:, None] * e2
    ang = rng.uniform(0, 2 * np.pi, n)
    b1 = e1 / np.maximum(np.linalg.norm(e1, axis=1), 1e-300)[:, None]
    b2 = np.cross(nrm, b1)
    din = np.cos(ang)[:, None] * b1 + np.sin(ang)[:, None] * b2
    detw = 10 ** rng.uniform(-5.05, -3, n) * rng.choice([-1, 1], n)
    dn = np.clip(detw / np.maximum(A2, 1e-300), -0.9, 0.9)
    d = din * np.sqrt(1 - dn ** 2)[:, None] + dn[:, None] * nrm
    d /= np.maximum(np.linalg.norm(d, axis=1), 1e-300)[:, None]
    with np.errstate(all="ignore"):
        back = np.where(d > 0, (P - lo) / d, (P - hi) / d)
    dist = np.min(np.where(np.isfinite(back) & (back > 0), back, np.inf), axis=1) * rng.uniform(0.02, 0.999, n)
    dist = np.where(np.isfinite(dist), dist, 0.5)
    o = (P - d * dist[:, None]).astype(f32)
    d = d.astype(f32)
    keep = np.all(np.isfinite(o), 1) & np.all(np.isfinite(d), 1) & np.all((o >= lo) & (o <= hi), 1) & (A2 > 0)
    return o[keep], d[keep], k[keep]
