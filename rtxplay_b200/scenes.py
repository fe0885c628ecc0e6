"""Synthetic scenes shared by the CUDA path and the oracle (one generator, one list).

SURVEY.md 7/H1: the reference's two programs do not place the same spheres (rtow.cxx:59
and optx/rtwo.cxx:167-172 consume rand() in different orders and precisions), so parity
needs ONE generator whose output list is handed to both sides.  book1() follows the
recipe of RTWO::load (optx/rtwo.cxx:137-244): ground, a (2*grid)^2 field of small
spheres with the 80/15/5 material split, three big spheres, and the reference's mesh
assignment (subdivisions 9 / 6 / 3 / 8 / 6 / 3), with float32 values drawn from a seeded
numpy stream.
"""
import numpy as np

from . import api

TH_KIND, TH_MESH, TH_XF, TH_TYPE, TH_ALB, TH_FUZZ, TH_INDEX, TH_STRIDE = 0, 1, 2, 14, 15, 18, 19, 20


def _sphere(center, radius, type, albedo=(0., 0., 0.), fuzz=0., index=0., ndiv=6):
    f = np.float32
    return dict(center=tuple(float(f(c)) for c in center), radius=float(f(radius)), type=int(type),
                albedo=tuple(float(f(a)) for a in albedo), fuzz=float(f(fuzz)), index=float(f(index)), ndiv=int(ndiv))


def book1(seed=1, grid=11):
    """RTOW book-1 final scene, optx/rtwo.cxx:137-244 (rtow.cxx:51-80): list of sphere dicts."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f = np.float32

    def rnd(lo=0., hi=1.):
        return f(lo) + f(rng.random(dtype=np.float32)) * (f(hi) - f(lo))

    s = [_sphere((0., -1000., 0.), 1000., api.DIFFUSE, (.5, .5, .5), ndiv=9)]
    for a in range(-grid, grid):
        for b in range(-grid, grid):
            select = rnd()
            center = (f(a) + f(.9) * rnd(), f(.2), f(b) + f(.9) * rnd())
            d = np.array(center, dtype=np.float32) - np.array((4., .2, 0.), dtype=np.float32)
            if np.sqrt(np.dot(d, d)) > f(.9):
                if select < f(.8):
                    alb = tuple(rnd() * rnd() for _ in range(3))
                    s.append(_sphere(center, .2, api.DIFFUSE, alb, ndiv=6))
                elif select < f(.95):
                    alb = tuple(rnd(.5, 1.) for _ in range(3))
                    s.append(_sphere(center, .2, api.REFLECT, alb, fuzz=rnd(0., .5), ndiv=6))
                else:
                    s.append(_sphere(center, .2, api.REFRACT, index=1.5, ndiv=3))
    s.append(_sphere((0., 1., 0.), 1., api.REFRACT, index=1.5, ndiv=8))
    s.append(_sphere((-4., 1., 0.), 1., api.DIFFUSE, (.4, .2, .1), ndiv=6))
    s.append(_sphere((4., 1., 0.), 1., api.REFLECT, (.7, .6, .5), fuzz=0., ndiv=3))
    return s


def grid_field(n_side, seed=2, ndiv=6):
    """Stress scene (BASELINE.json config 5): n_side^2 small spheres on a square grid over
    the ground sphere, same material mix."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f = np.float32
    s = [_sphere((0., -1000., 0.), 1000., api.DIFFUSE, (.5, .5, .5), ndiv=9)]
    half = n_side // 2
    for a in range(-half, n_side - half):
        for b in range(-half, n_side - half):
            u = rng.random(6, dtype=np.float32)
            center = (f(a) * f(.5) + f(.25) * u[0], f(.1), f(b) * f(.5) + f(.25) * u[1])
            if u[2] < .8:
                s.append(_sphere(center, .1, api.DIFFUSE, (u[3] * u[3], u[4] * u[4], u[5] * u[5]), ndiv=ndiv))
            elif u[2] < .95:
                s.append(_sphere(center, .1, api.REFLECT, (.5 + .5 * u[3], .5 + .5 * u[4], .5 + .5 * u[5]), fuzz=.5 * u[0], ndiv=ndiv))
            else:
                s.append(_sphere(center, .1, api.REFRACT, index=1.5, ndiv=ndiv))
    return s


def xf_of(sp):
    """Row-major 3x4 transform of a sphere instance: uniform scale + translation
    (optx/rtwo.cxx:159-165)."""
    if "xf" in sp:                      # a general affine instance transform
        return np.asarray(sp["xf"], dtype=np.float32).reshape(12)
    r, (cx, cy, cz) = sp["radius"], sp["center"]
    return np.array([r, 0, 0, cx, 0, r, 0, cy, 0, 0, r, cz], dtype=np.float32)


def affine_mix(seed=7, n=24):
    """A small scene whose mesh instances carry general affine transforms (rotation, non-uniform
    scale, shear) -- exercises the non-diagonal transform path; mesh mode only."""
    rng = np.random.Generator(np.random.PCG64(seed))
    s = [_sphere((0., -1000., 0.), 1000., api.DIFFUSE, (.5, .5, .5), ndiv=3)]
    for k in range(n):
        c = np.array([rng.uniform(-4, 4), rng.uniform(.4, 1.5), rng.uniform(-4, 4)])
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        ang = rng.uniform(0, np.pi)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        S = np.diag(rng.uniform(.15, .5, 3))
        H = np.eye(3)
        H[0, 1] = rng.uniform(-.3, .3)
        M = R @ S @ H
        t = int(rng.integers(0, 3))
        sp = _sphere(tuple(c), .3, t, tuple(rng.uniform(.2, .9, 3)), fuzz=rng.uniform(0, .4), index=1.5, ndiv=2)
        sp["xf"] = np.concatenate([M, c[:, None]], axis=1).astype(np.float32).reshape(12)
        s.append(sp)
    return s


def table(spheres, mode="analytic", ndiv=None):
    """The thing table the oracle reads (oracle/oracle.cxx "scene tables") and the meshes it
    references.  mode 'mesh': every sphere instances the unit-sphere mesh of its subdivision
    count (or of `ndiv` for all, to keep brute-force oracles affordable)."""
    t = np.zeros((len(spheres), TH_STRIDE), dtype=np.float64)
    meshes, mesh_of = [], {}
    for k, sp in enumerate(spheres):
        t[k, TH_XF:TH_XF + 12] = xf_of(sp)
        t[k, TH_TYPE] = sp["type"]
        t[k, TH_ALB:TH_ALB + 3] = sp["albedo"]
        t[k, TH_FUZZ], t[k, TH_INDEX] = sp["fuzz"], sp["index"]
        if mode == "analytic":
            t[k, TH_KIND], t[k, TH_MESH] = 0, -1
        else:
            n = sp["ndiv"] if ndiv is None else ndiv
            if n not in mesh_of:
                mesh_of[n] = len(meshes)
                meshes.append(api.sphere_mesh(1., n))
            t[k, TH_KIND], t[k, TH_MESH] = 1, mesh_of[n]
    return t, meshes


def load(ctx, spheres, mode="analytic", ndiv=None):
    """RTWO::load() + Scene::build (optx/rtwo.cxx:137-250) on a Context: creates the meshes,
    adds one thing per sphere in list order (thing id = list index), builds the top level.
    Returns (table, meshes) for the oracle."""
    t, meshes = table(spheres, mode, ndiv)
    if mode == "analytic":
        ids = {-1: ctx.add_analytic_sphere()}
    else:
        ids = {q: ctx.add_mesh(v, i) for q, (v, i) in enumerate(meshes)}
    for k, sp in enumerate(spheres):
        tid = ctx.add_thing(ids[int(t[k, TH_MESH])], api.Optics(sp["type"], sp["albedo"], sp["fuzz"], sp["index"]), xf_of(sp))
        assert tid == k
    ctx.build()
    return t, meshes
