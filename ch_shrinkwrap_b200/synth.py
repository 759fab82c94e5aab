"""Seeded synthetic localisation clouds and starting meshes (measurement input).

Restates, with fixed seeds, the pieces of the reference's simulation path that the
BASELINE configs name: signed-distance primitives (``sdf.py:60-146``), the smooth
union used for necked shapes (``shape.py:368-376``), the exponential-photon
localisation-precision model (``util.py:37-47``, defaults ``simulation.py:19-24``) and the
uniform background model (``evaluation_utils.py:230-256``).  ``Shape.points()``
itself needs PYME's ``points_from_sdf`` (``shape.py:16,75``), which is absent, so surface
samples are drawn here by projecting jittered seeds onto the zero level set.
"""
from __future__ import annotations

import numpy as np

from .minimesh import MiniMesh, geodesic_sphere


# ---- signed distance functions (points as (N,3)) ----------------------------------
class Sphere:
    def __init__(self, radius=500.0, centre=(0.0, 0.0, 0.0)):
        self.radius = float(radius)
        self.centre = np.asarray(centre, dtype=np.float64)
        self.bound = self.radius + float(np.abs(self.centre).max())

    def sdf(self, p):
        return np.sqrt(((p - self.centre) ** 2).sum(1)) - self.radius


class Ellipsoid:
    """Bound-type ellipsoid distance (first-order accurate near the surface)."""

    def __init__(self, a=600.0, b=400.0, c=300.0):
        self.r = np.array([a, b, c], dtype=np.float64)
        self.bound = float(self.r.max())

    def sdf(self, p):
        k0 = np.sqrt(((p / self.r) ** 2).sum(1))
        k1 = np.sqrt(((p / (self.r * self.r)) ** 2).sum(1))
        with np.errstate(invalid='ignore', divide='ignore'):
            return np.where(k1 > 0, k0 * (k0 - 1.0) / k1, -self.r.min())


class Capsule:
    def __init__(self, start=(-300.0, 0.0, 0.0), end=(300.0, 0.0, 0.0), radius=200.0):
        self.a = np.asarray(start, dtype=np.float64)
        self.b = np.asarray(end, dtype=np.float64)
        self.radius = float(radius)
        self.bound = float(max(np.abs(self.a).max(), np.abs(self.b).max()) + self.radius)

    def sdf(self, p):
        pa = p - self.a
        ba = self.b - self.a
        h = np.clip((pa @ ba) / (ba @ ba), 0.0, 1.0)
        d = pa - h[:, None] * ba
        return np.sqrt((d * d).sum(1)) - self.radius


class SmoothUnion:
    """min(d0,d1) - h^2/(4k), h = max(k-|d0-d1|,0)   (shape.py:368-376)."""

    def __init__(self, s0, s1, k=0.0):
        self.s0, self.s1, self.k = s0, s1, float(k)
        self.bound = max(s0.bound, s1.bound)

    def sdf(self, p):
        d0, d1 = self.s0.sdf(p), self.s1.sdf(p)
        res = np.minimum(d0, d1)
        if self.k > 0:
            h = np.maximum(self.k - np.abs(d0 - d1), 0.0)
            res = res - h * h * 0.25 / self.k
        return res


def two_lobed(radius=400.0, sep=560.0, k=120.0):
    """Necked dumbbell: smooth union of two spheres (BASELINE config 3)."""
    return SmoothUnion(Sphere(radius, (-sep / 2, 0, 0)), Sphere(radius, (sep / 2, 0, 0)), k)


def sdf_grad(shape, p, h=0.5):
    g = np.empty_like(p)
    for k in range(3):
        e = np.zeros(3)
        e[k] = h
        g[:, k] = (shape.sdf(p + e) - shape.sdf(p - e)) / (2 * h)
    return g


def project_to_surface(shape, p, n_steps=6):
    """Newton projection p <- p - sdf * grad/|grad|^2 onto the zero level set."""
    p = np.array(p, dtype=np.float64)
    for _ in range(n_steps):
        d = shape.sdf(p)
        g = sdf_grad(shape, p)
        gg = np.maximum((g * g).sum(1), 1e-12)
        p -= (d / gg)[:, None] * g
    return p


def radial_surface(shape, dirs, r_max=None, n_bisect=40):
    """Star-shaped shapes: distance along unit ``dirs`` from the origin to the surface."""
    r_max = 2.0 * shape.bound if r_max is None else r_max
    lo = np.zeros(len(dirs))
    hi = np.full(len(dirs), float(r_max))
    for _ in range(n_bisect):
        mid = 0.5 * (lo + hi)
        inside = shape.sdf(dirs * mid[:, None]) < 0
        lo = np.where(inside, mid, lo)
        hi = np.where(inside, hi, mid)
    return 0.5 * (lo + hi)


# ---- noise model ---------------------------------------------------------------
def loc_error(n, rng, psf_width=(280.0, 280.0, 840.0), mean_photon_count=600.0, bg_photon_count=20.0):
    """Per-axis localisation precision, exponential photon model (util.py:37-47)."""
    sig = np.empty((n, 3), dtype=np.float64)
    for k in range(3):
        out = np.empty(0)
        while len(out) < n:
            l = rng.exponential(mean_photon_count, 2 * (n - len(out)) + 16)
            out = np.concatenate([out, l[l > bg_photon_count]])
        sig[:, k] = (psf_width[k] / 2.355) / np.sqrt(out[:n])
    return sig


def smlm_cloud(shape, n_points, seed=0, noise_fraction=0.1, dtype=np.float32, star=True,
               psf_width=(280.0, 280.0, 840.0), mean_photon_count=600.0, bg_photon_count=20.0):
    """Noisy localisations on ``shape`` plus uniform background.

    Returns (points (P,3), sigma (P,3)) in nm.  ``noise_fraction`` of the points are
    uniform over 1.2x the bounding box (evaluation_utils.py:230-243).
    """
    rng = np.random.default_rng(seed)
    n_bg = int(round(n_points * noise_fraction))
    n_s = n_points - n_bg
    d = rng.standard_normal((n_s, 3))
    d /= np.sqrt((d * d).sum(1))[:, None]
    if star:
        r = radial_surface(shape, d)
        p = d * r[:, None]
    else:
        p = project_to_surface(shape, d * shape.bound * rng.uniform(0.3, 1.0, n_s)[:, None])
    sig = loc_error(n_points, rng, psf_width, mean_photon_count, bg_photon_count)
    p = p + sig[:n_s] * rng.standard_normal((n_s, 3))
    lo, hi = p.min(0), p.max(0)
    c, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * 1.2
    bg = c + half * rng.uniform(-1.0, 1.0, (n_bg, 3))
    pts = np.concatenate([p, bg], 0)
    perm = rng.permutation(n_points)
    return np.ascontiguousarray(pts[perm], dtype=dtype), np.ascontiguousarray(sig[perm], dtype=dtype)


def fast_sphere_cloud(n_points, radius=500.0, seed=0, noise_fraction=0.1, chunk=1 << 22):
    """Large float32 sphere clouds for the bench, generated in chunks (same noise model)."""
    rng = np.random.default_rng(seed)
    pts = np.empty((n_points, 3), np.float32)
    sig = np.empty((n_points, 3), np.float32)
    psf = np.array([280.0, 280.0, 840.0], np.float32) / np.float32(2.355)
    for s in range(0, n_points, chunk):
        e = min(n_points, s + chunk)
        m = e - s
        d = rng.standard_normal((m, 3), dtype=np.float32)
        d /= np.sqrt((d * d).sum(1))[:, None]
        l = rng.exponential(600.0, (m, 3)).astype(np.float32)
        l = np.maximum(l, np.float32(20.0))  # clamp instead of reject: keeps chunks independent
        sg = psf / np.sqrt(l)
        p = d * np.float32(radius) + sg * rng.standard_normal((m, 3), dtype=np.float32)
        bgm = rng.random(m) < noise_fraction
        nb = int(bgm.sum())
        p[bgm] = (rng.random((nb, 3), dtype=np.float32) * 2 - 1) * np.float32(1.2 * radius)
        pts[s:e] = p
        sig[s:e] = sg
    return pts, sig


# ---- starting meshes -----------------------------------------------------------
def star_mesh(shape, n, scale=1.2):
    """Geodesic sphere radially projected onto ``scale`` x the (star-shaped) shape.

    Stands in for the coarse isosurface the recipe starts from (surface_fitting.py:56).
    """
    v, f = geodesic_sphere(n)
    r = radial_surface(shape, v)
    return MiniMesh(v * (scale * r)[:, None], f)


def mesh_surface_cloud(verts, faces, n_points, seed=0, noise_fraction=0.1, chunk=1 << 20, threads=None,
                       psf_width=(280.0, 280.0, 840.0), mean_photon_count=600.0, bg_photon_count=20.0):
    """Large float32 clouds for the bench: uniform samples on a fine triangulation of the shape, jittered with
    the exponential-photon precision model, plus uniform background.  Chunked with one child seed per chunk,
    so the result does not depend on the number of worker threads."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    verts = np.asarray(verts, dtype=np.float32)
    faces = np.asarray(faces)
    v0, v1, v2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    area = 0.5 * np.sqrt((np.cross((v1 - v0).astype(np.float64), (v2 - v0).astype(np.float64)) ** 2).sum(1))
    cdf = np.cumsum(area)
    cdf /= cdf[-1]
    lo, hi = verts.min(0), verts.max(0)
    c, half = 0.5 * (lo + hi), 0.5 * (hi - lo) * 1.2
    psf = (np.asarray(psf_width, np.float32) / np.float32(2.355))
    pts = np.empty((n_points, 3), np.float32)
    sig = np.empty((n_points, 3), np.float32)
    starts = list(range(0, n_points, chunk))
    seeds = np.random.SeedSequence(seed).spawn(len(starts))

    def work(k):
        s = starts[k]
        e = min(n_points, s + chunk)
        m = e - s
        rng = np.random.default_rng(seeds[k])
        f = np.minimum(np.searchsorted(cdf, rng.random(m)), len(faces) - 1)
        r1 = np.sqrt(rng.random(m, dtype=np.float32))
        r2 = rng.random(m, dtype=np.float32)
        a, b = (1 - r1)[:, None], (r1 * (1 - r2))[:, None]
        p = a * v0[f] + b * v1[f] + (1 - a - b) * v2[f]
        l = rng.exponential(mean_photon_count, (m, 3)).astype(np.float32)
        l = np.maximum(l, np.float32(bg_photon_count))       # clamp instead of reject: keeps chunks independent
        sg = psf / np.sqrt(l)
        p += sg * rng.standard_normal((m, 3), dtype=np.float32)
        bgm = rng.random(m) < noise_fraction
        nb = int(bgm.sum())
        p[bgm] = c + half * (rng.random((nb, 3), dtype=np.float32) * 2 - 1)
        pts[s:e] = p
        sig[s:e] = sg

    with ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 8)) as ex:
        list(ex.map(work, range(len(starts))))
    return pts, sig
