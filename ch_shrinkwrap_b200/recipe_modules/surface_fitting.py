"""``ShrinkwrapMembrane`` recipe module with the reference's parameters and ``execute`` contract
(``ch_shrinkwrap/recipe_modules/surface_fitting.py:11-115``), driving the GPU path.

With PYME installed the class is a real ``ModuleBase`` with traits and registers itself as
``ShrinkwrapMembrane``; without PYME (this container) it is a plain object with the same attribute names and
defaults, usable on any dict-like namespace whose ``surf`` entry exposes ``vertices``/``faces`` (or PYME's
``_vertices``/``faces``) and whose ``points`` entry maps column names to arrays.
"""
from __future__ import annotations

import logging
import time

import numpy as np

logger = logging.getLogger(__name__)

_DEFAULTS = dict(
    input='surf', output='membrane', points='filtered_localizations',
    max_iters=39, curvature_weight=20.0, finishing_iters=0, finishing_curvature_weight=20.0, shrink_weight=0.0,
    kc=1.0, remesh_frequency=5, punch_frequency=0, min_hole_radius=100.0,
    sigma_x='error_x', sigma_y='error_y', sigma_z='error_z',
    neck_threshold_low=-1e-3, neck_threshold_high=1e-2, neck_first_iter=9, truncate_at=1000,
    minimum_edge_length=5.0, smooth_curvature=True)

try:                                                       # pragma: no cover - PYME is not in this image
    from PYME.recipes.base import ModuleBase, register_module
    from PYME.recipes.traits import Bool, CStr, Float, Input, Int, Output
    _HAVE_PYME = True
except Exception:                                          # noqa: BLE001
    _HAVE_PYME = False

    class ModuleBase(object):
        def __init__(self, **kwargs):
            for k, v in _DEFAULTS.items():
                setattr(self, k, v)
            for k, v in kwargs.items():
                if k not in _DEFAULTS:
                    raise TypeError('unknown parameter %r' % k)
                setattr(self, k, v)

    def register_module(name):
        return lambda cls: cls


_GPU_MESH_CLASS = {}


def gpu_mesh_class(host_cls):
    """The host mesh class with the GPU NanoWrap path mixed in FRONT of it: ``opt_conjugate_gradient``, ``shrink_wrap``,
    the curvature properties and ``remove_necks`` resolve to ``ShrinkwrapMeshMixin`` (libnanowrap.so), everything
    topological (remesh, repair, punch_holes, ...) to the host class.  Without this, ``_membrane_mesh.MembraneMesh`` would
    run its own ``opt_conjugate_gradient``, which builds the reference's CPU ``ShrinkwrapMeshConjGrad``
    (_membrane_mesh.pyx:1428,1510) -- a silent CPU path."""
    cls = _GPU_MESH_CLASS.get(host_cls)
    if cls is None:
        from ..membrane_mesh import ShrinkwrapMeshMixin
        cls = type('GpuMembraneMesh', (ShrinkwrapMeshMixin, host_cls), {'__doc__': gpu_mesh_class.__doc__})
        _GPU_MESH_CLASS[host_cls] = cls
    return cls


def _mesh_factory(inp, **kw):
    """MembraneMesh(mesh=inp, ...) as at surface_fitting.py:56: the reference's (PYME-backed) mesh class with the GPU path
    mixed in when it is importable, else the harness mesh.  Either way the solver that runs is the GPU one."""
    try:
        from ch_shrinkwrap import _membrane_mesh
        host_cls = _membrane_mesh.MembraneMesh
    except Exception:                                      # noqa: BLE001  (no PYME / reference extension in this environment)
        from ..membrane_mesh import MembraneMesh
        return MembraneMesh(mesh=inp, **kw)
    mesh = gpu_mesh_class(host_cls)(mesh=inp, **kw)
    if not hasattr(mesh, 'cg'):
        mesh.cg = None
    return mesh


@register_module('ShrinkwrapMembrane')
class ShrinkwrapMembrane(ModuleBase):
    if _HAVE_PYME:                                         # pragma: no cover
        input = Input('surf'); output = Output('membrane'); points = Input('filtered_localizations')
        max_iters = Int(39); curvature_weight = Float(20.0); finishing_iters = Int(0)
        finishing_curvature_weight = Float(20.0); shrink_weight = Float(0); kc = Float(1.0)
        remesh_frequency = Int(5); punch_frequency = Int(0); min_hole_radius = Float(100.0)
        sigma_x = CStr('error_x'); sigma_y = CStr('error_y'); sigma_z = CStr('error_z')
        neck_threshold_low = Float(-1e-3); neck_threshold_high = Float(1e-2); neck_first_iter = Int(9)
        truncate_at = Int(1000); minimum_edge_length = Float(5); smooth_curvature = Bool(True)

    def execute(self, namespace):
        inp = namespace[self.input]
        n_faces = len(inp.faces)
        if not n_faces > 4:
            raise RuntimeError('Input mesh only has %d faces, a valid surface needs at least 4 faces' % n_faces)
        mesh = _mesh_factory(inp, kc=self.kc, max_iter=self.max_iters, step_size=self.curvature_weight,
                             remesh_frequency=self.remesh_frequency, delaunay_remesh_frequency=self.punch_frequency,
                             delaunay_eps=self.min_hole_radius, neck_threshold_low=self.neck_threshold_low,
                             neck_threshold_high=self.neck_threshold_high, neck_first_iter=self.neck_first_iter,
                             shrink_weight=self.shrink_weight, truncate_at=self.truncate_at)
        namespace[self.output] = mesh
        src = namespace[self.points]
        pts = np.ascontiguousarray(np.vstack([src['x'], src['y'], src['z']]).T)
        try:
            sigma = np.vstack([src[self.sigma_x], src[self.sigma_y], src[self.sigma_z]]).T
        except Exception:                                  # noqa: BLE001  (same fallbacks as surface_fitting.py:79-92)
            try:
                sigma = src[self.sigma_x]
            except KeyError:
                print(f"{self.sigma_x} not found in data source, defaulting to 10 nm precision.")
                sigma = 10 * np.ones_like(src['x'])
        start = time.time()
        mesh.shrink_wrap(pts, sigma, method='conjugate_gradient', minimum_edge_length=self.minimum_edge_length)
        if self.finishing_iters > 0:
            mesh.step_size = self.finishing_curvature_weight
            mesh.shrink_wrap(pts, sigma, method='conjugate_gradient', minimum_edge_length=self.minimum_edge_length,
                             max_iter=self.finishing_iters)
        if self.smooth_curvature:
            mesh.smooth_curvature = self.smooth_curvature
            mesh._populate_curvature_grad()
        duration = time.time() - start
        md = dict(getattr(inp, 'mdh', None) or {})
        md['Processing.ShrinkwrapMembrane.Runtime'] = duration          # surface_fitting.py:110
        for k in _DEFAULTS:
            md['Processing.ShrinkwrapMembrane.' + k] = getattr(self, k)
        mesh.mdh = md
        return mesh


_IMAGE_DEFAULTS = dict(
    input='surf', output='membrane', input_image='input',
    max_iters=100, curvature_weight=10.0, shrink_weight=1.0, kc=1.0, remesh_frequency=5, cut_frequency=0,
    min_hole_radius=100.0, sigma_x='sigma_x', sigma_y='sigma_y', sigma_z='sigma_z',
    neck_threshold_low=-1e-4, neck_threshold_high=1e-2, neck_first_iter=9, minimum_edge_length=-1.0)


def image_to_weighted_points(data, voxelsize_nm, origin):
    """Voxels with positive intensity as weighted localisations (recipe_modules/surface_fitting.py:305-327):
    returns (pts (P,3) float64, weights (3P,), sigma scalar = vx)."""
    weights = np.asarray(data)
    vx, vy, vz = voxelsize_nm
    ox, oy, oz = origin
    x, y, z = np.mgrid[0:weights.shape[0], 0:weights.shape[1], 0:weights.shape[2]]
    x = ox + vx * x.ravel()
    y = oy + vy * y.ravel()
    z = oz + vz * z.ravel()
    weights = weights.ravel()
    mask = weights > 0
    weights = weights[mask]
    pts = np.ascontiguousarray(np.vstack([x[mask], y[mask], z[mask]]).T)
    return pts, np.repeat(weights, 3), vx


@register_module('ImageShrinkwrapMembrane')
class ImageShrinkwrapMembrane(ModuleBase):
    """Same solver driven by an image: every voxel with positive intensity is a point, its intensity the weight, the
    voxel size the (scalar, un-inverted) sigma -- reference: recipe_modules/surface_fitting.py:246-341."""
    if _HAVE_PYME:                                         # pragma: no cover
        input = Input('surf'); output = Output('membrane'); input_image = Input('input')
        max_iters = Int(100); curvature_weight = Float(10.0); shrink_weight = Float(1.0); kc = Float(1.0)
        remesh_frequency = Int(5); cut_frequency = Int(0); min_hole_radius = Float(100.0)
        sigma_x = CStr('sigma_x'); sigma_y = CStr('sigma_y'); sigma_z = CStr('sigma_z')
        neck_threshold_low = Float(-1e-4); neck_threshold_high = Float(1e-2); neck_first_iter = Int(9)
        minimum_edge_length = Float(-1.0)
    else:
        def __init__(self, **kwargs):
            for k, v in _IMAGE_DEFAULTS.items():
                setattr(self, k, v)
            for k, v in kwargs.items():
                if k not in _IMAGE_DEFAULTS:
                    raise TypeError('unknown parameter %r' % k)
                setattr(self, k, v)

    def execute(self, namespace):
        inp = namespace[self.input]
        n_faces = len(inp.faces)
        if not n_faces > 4:
            raise RuntimeError('Input mesh only has %d faces, a valid surface needs at least 4 faces' % n_faces)
        mesh = _mesh_factory(inp, kc=self.kc, max_iter=self.max_iters, step_size=self.curvature_weight,
                             remesh_frequency=self.remesh_frequency, delaunay_remesh_frequency=self.cut_frequency,
                             delaunay_eps=self.min_hole_radius, neck_threshold_low=self.neck_threshold_low,
                             neck_threshold_high=self.neck_threshold_high, neck_first_iter=self.neck_first_iter,
                             shrink_weight=self.shrink_weight)
        if hasattr(mesh, 'repair'):                        # host topology (PYME): close holes before starting (:291-293)
            mesh.repair()
            mesh.remesh()
        namespace[self.output] = mesh
        im = namespace[self.input_image]
        data = im.data_xyztc[:, :, :, 0, 0] if hasattr(im, 'data_xyztc') else np.asarray(im.data)
        pts, weights, sigma = image_to_weighted_points(data, im.voxelsize_nm, im.origin)
        mesh.shrink_wrap(pts, sigma=sigma, weights=weights, method='conjugate_gradient',
                         minimum_edge_length=self.minimum_edge_length)
        md = dict(getattr(inp, 'mdh', None) or {})
        for k in _IMAGE_DEFAULTS:
            md['Processing.ImageShrinkwrapMembrane.' + k] = getattr(self, k)
        mesh.mdh = md
        return mesh
