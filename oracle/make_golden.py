"""Writes tests/golden/*.npz by running the UNMODIFIED reference (this container only; needs /root/reference).

    python -m oracle.make_golden

Each fixture stores the seeded inputs' parameters (the inputs themselves are regenerated from the seed by
tests/conftest.make_case and cross-checked by a checksum) and the reference's outputs:
  search_sphere_f32 : ShrinkwrapMeshConjGrad.search, 6 iterations (mesh_conj_grad.py:150-292): nearest faces and
                      weights of the last iteration, residual, S, histories, final vertices
  search_lobed_f64  : same on the two-lobed shape with float64 points and a scalar (un-inverted) sigma
  curvature_sphere  : c_curvature_grad (membrane_mesh_utils.c:915) after srand(42), plus the rand() stream used
  ring_ops          : c_shrinkwrap_l/lh/lw/lhw_func and vertex_area_weights (conj_grad_utils.c:249-710)
"""
from __future__ import annotations

import copy
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from conftest import make_case  # noqa: E402
from oracle import refharness  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest()[:8], dtype=np.uint64)[0]


def search_fixture(name, case_kw, sigma_mode, lam, n_iters):
    from ch_shrinkwrap_b200 import synth
    kw = dict(case_kw)
    if kw.pop('lobed', False):
        kw['shape'] = synth.two_lobed()
    mesh, pts, sig = make_case(**kw)
    if sigma_mode == 'array':
        s = (1.0 / sig.ravel()).astype(pts.dtype)
    else:
        s = 10.0
    cg = refharness.reference_solver(mesh, pts)
    v = cg.search(pts, lams=[lam], num_iters=n_iters, sigma_inv=s)
    v_idx, w = cg.w
    np.savez_compressed(
        os.path.join(OUT, name + '.npz'),
        input_digest=digest(pts, sig, mesh.faces), lam=lam, n_iters=n_iters, sigma_mode=sigma_mode,
        v_idx=v_idx.astype(np.int32), w=w, d=cg.d[:, 0], res=np.asarray(cg.res), S=cg.S,
        tests=np.array(cg.tests, np.float64), ress=np.array(cg.ress, np.float64),
        prefs=np.array([float(p[0]) for p in cg.prefs]), vertices=v, cpred=float(cg.cpred), wpred=float(cg.wpreds[0]))
    print(name, 'ok', v.shape, cg.tests[-1])


def main():
    os.makedirs(OUT, exist_ok=True)
    search_fixture('search_sphere_f32', dict(n_points=3000, n_geo=5, seed=101), 'array', 10.0, 6)
    search_fixture('search_lobed_f64', dict(n_points=2500, n_geo=6, seed=102, dtype=np.float64, lobed=True), 'scalar', 5.0, 4)

    from ch_shrinkwrap_b200 import minimesh
    m = minimesh.sphere_mesh(50.0, 6)
    nv = int((m._vertices['halfedge'] != -1).sum())
    u = refharness.libc_uniforms(42, 3 * nv)
    r = refharness.reference_curvature(m, seed=42)
    np.savez_compressed(os.path.join(OUT, 'curvature_sphere.npz'), jitter_u=u, radius=50.0, n_geo=6, **r)
    print('curvature_sphere ok', float(r['H'].mean()) * 50)

    _, cgu = refharness.load_reference()
    mesh, pts, _ = make_case(n_points=50, n_geo=4, seed=103)
    nb = mesh.neighbor_vertices()
    M = nb.shape[0]
    rng = np.random.default_rng(7)
    f = rng.standard_normal(3 * M).astype(np.float32)
    ref = np.ascontiguousarray(mesh.vertices, dtype=np.float32).ravel()
    out = {}
    for key, fn, third in (('l', cgu.c_shrinkwrap_l_func, None), ('lh', cgu.c_shrinkwrap_lh_func, None),
                           ('lw', cgu.c_shrinkwrap_lw_func, ref), ('lhw', cgu.c_shrinkwrap_lhw_func, ref)):
        d = np.zeros(3 * M, np.float32)
        fn(f, nb, d if third is None else third, d, 3, 50, M, nb.shape[1])
        out[key] = d
    w = np.zeros(3 * M, np.float32)
    cgu.vertex_area_weights(ref, nb, w, M, nb.shape[1])
    out['area_w'] = w
    np.savez_compressed(os.path.join(OUT, 'ring_ops.npz'), f=f, seed=103, **out)
    print('ring_ops ok')
    quality_fixture()


def quality_fixture():
    """Quality metrics and hole-punch searches from the unmodified reference:
    evaluation_utils.points_from_mesh / average_squared_distance (imported), c_holepunch_pair_candidate_faces (compiled
    from the reference's source through oracle/ref_curvature_wrapper.c)."""
    from ch_shrinkwrap_b200 import synth
    evu = refharness.load_evaluation_utils()
    shape = synth.two_lobed()
    m = synth.star_mesh(shape, 4, scale=1.0)
    np.random.seed(5)
    d = evu.points_from_mesh(m, dx_min=25)                      # random order (np.random.choice, :137)
    order = np.lexsort((d[:, 2], d[:, 1], d[:, 0]))
    pts, _ = synth.smlm_cloud(shape, 4000, seed=104)
    msd = evu.average_squared_distance(d, pts.astype(np.float64))
    msd32 = evu.average_squared_distance(d.astype(np.float32), pts)
    cand = np.arange(len(m._faces), dtype=np.int32)[::2].copy()
    pairs = refharness.reference_holepunch_pairs(m, cand)
    np.savez_compressed(os.path.join(OUT, 'quality.npz'), n_geo=4, dx_min=25, samples_sorted=d[order], cloud_seed=104, cloud_n=4000,
                        msd=np.array(msd), msd32=np.array(msd32), candidates=cand, pairs=pairs)
    print('quality ok', d.shape, msd, int((pairs != -1).sum()))


if __name__ == '__main__':
    main()
