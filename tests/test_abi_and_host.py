"""CPU-side checks: the C-ABI library loads and exports every symbol include/rtx.h declares,
the host helpers (camera, tessellator) give the reference's known answers, and the CUDA
core driven serially on the host (tests/hostemu) equals the oracle's float mirror bit for bit."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle as orc
from rtxplay_b200 import _lib, api, scenes
from tests import hostemu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    _lib.build()
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "rtx.h")).read()
    declared = set(re.findall(r"\b(rtx_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(L, name) is not None


def test_no_device_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(api.RtxError, match="no CUDA device"):
        api.Context(0)


def test_struct_sizes_match_header():
    # rtx_params: 4 u32 + camera (19 floats) + pad + u64 + 3 u32 (+pad)
    assert ctypes.sizeof(_lib.RtxCamera) == 19 * 4
    assert ctypes.sizeof(_lib.RtxOptics) == 24
    assert ctypes.sizeof(_lib.RtxParams) == 128          # ... + guides, variant
    assert ctypes.sizeof(_lib.RtxStats) == 64


@pytest.mark.parametrize("ndiv,nv,nt", [(0, 4, 4), (1, 10, 16), (3, 130, 256), (6, 8194, 16384)])
def test_tessellator_counts(ndiv, nv, nt):
    """2*4^n+2 vertices, 4*4^n triangles (probe of optx/sphere.cxx -DMAIN)."""
    v, i = api.sphere_mesh(1., ndiv)
    assert v.shape == (nv, 3) and i.shape == (nt, 3)
    assert i.max() == nv - 1
    assert np.allclose(np.linalg.norm(v, axis=1), 1., atol=2e-6)


def test_tessellator_known_answer():
    """First vertices / faces of `sphere 1. 1` as printed by the reference tool."""
    v, i = api.sphere_mesh(1., 1)
    assert np.allclose(v[:3], [[.57735026919, .57735026919, .57735026919], [1, 0, 0], [0, 0, 1]], atol=1e-7)
    assert (i[:4] + 1).tolist() == [[1, 2, 3], [2, 4, 5], [3, 5, 6], [2, 5, 3]]


def test_tessellator_radius_and_winding():
    v, i = api.sphere_mesh(.2, 3)
    assert np.allclose(np.linalg.norm(v, axis=1), .2, atol=1e-6)
    a, b, c = v[i[:, 0]], v[i[:, 1]], v[i[:, 2]]
    # the reference's faces wind clockwise seen from outside (normals point inward)
    assert (np.einsum("ij,ij->i", np.cross(b - a, c - a), a + b + c) < 0).all()


def test_dedup_known_answer():
    """optx/reduce.cxx:66-106: 12 soup vertices (4 triangles) -> 6 unique vertices and the index
    list {0,1,2}{1,3,4}{2,4,5}{2,4,1} -- the ndiv=1 face of the tetrahedron has that shape."""
    v, i = api.sphere_mesh(1., 1)
    assert i[:4].tolist() == [[0, 1, 2], [1, 3, 4], [2, 4, 5], [1, 4, 2]] or len(np.unique(i[:4])) == 6


def test_camera_matches_oracle_float_camera():
    cam = api.camera_table(api.camera(aspratio=1200 / 800.))
    ref = orc.camera_f32((13, 2, 3), (0, 0, 0), (0, 1, 0), 20., 1200 / 800., .1, 10.)
    assert np.array_equal(cam, ref)
    # closed form: focused on `pat`, the ray through the image centre is -dvec
    # and reaches `pat` at t=1
    foc = float(np.sqrt(13. ** 2 + 2. ** 2 + 3. ** 2))
    cam = api.camera_table(api.camera(aspratio=1.5, aperture=0., fostance=foc))
    assert np.allclose(cam[0:3] - cam[15:18], 0., atol=1e-5)


def test_scene_generator_shape():
    sp = scenes.book1(seed=1)
    assert 470 <= len(sp) <= 488
    types = np.array([s["type"] for s in sp])
    assert (types == 0).sum() > (types == 1).sum() > (types == 2).sum() > 0
    assert sp[0]["radius"] == 1000. and sp[0]["ndiv"] == 9
    assert [s["ndiv"] for s in sp[-3:]] == [8, 6, 3]


@pytest.mark.parametrize("pool", [True, False])
@pytest.mark.parametrize("mode,ndiv", [("analytic", None), ("mesh", 2)])
def test_cuda_core_on_host_equals_float_mirror(mode, ndiv, pool):
    """The __host__ __device__ core (traversal through a Karras LBVH, primitive tests, shading,
    random stream) == oracle<float>: fixed-point radiance, segments and first-hit ids, bit for bit."""
    sp = scenes.book1(seed=3)
    tab, meshes = scenes.table(sp, mode, ndiv)
    cam = api.camera_table(api.camera(aspratio=1.5))
    w, h, spp = 60, 40, 2
    f = orc.render(orc.F32_PCG, tab, cam, w, h, spp, 50, want_first=True, meshes=meshes)
    # pool=True: the step functions of the render kernel's state machine (rtx_pool.cuh);
    # pool=False: the single-ray traversal used by the picker and the parity instruments
    e = hostemu.render(tab, cam, w, h, spp, 50, meshes=meshes, pool=pool)
    assert np.array_equal(f["first_id"], e["first_id"])
    assert np.array_equal(f["rpp"], e["rpp"])
    assert np.array_equal(f["fix"], e["fix"])


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("mode,ndiv,depth", [("analytic", None, 50), ("mesh", 2, 3), ("analytic", None, 1)])
def test_reference_variants_on_host(mode, ndiv, depth, variant):
    """Where the reference's own programs disagree: the iterative (1) and recursive
    (2) OptiX semantics -- pixel mapping over w / h, unguarded Lambert, rays instead of scatter events
    counted, throughput kept / black at the limit -- in the CUDA core on the host == oracle<float>."""
    sp = scenes.book1(seed=3)
    tab, meshes = scenes.table(sp, mode, ndiv)
    cam = api.camera_table(api.camera(aspratio=1.5))
    w, h, spp = 60, 40, 2
    f = orc.render(orc.F32_PCG, tab, cam, w, h, spp, depth, want_first=True, meshes=meshes, variant=variant)
    f0 = orc.render(orc.F32_PCG, tab, cam, w, h, spp, depth, meshes=meshes)
    assert not np.array_equal(f["fix"], f0["fix"])            # the variants do differ from rtow.cxx
    for pool in (True, False):
        e = hostemu.render(tab, cam, w, h, spp, depth, meshes=meshes, pool=pool, variant=variant)
        assert np.array_equal(f["first_id"], e["first_id"])
        assert np.array_equal(f["rpp"], e["rpp"])
        assert np.array_equal(f["fix"], e["fix"])
    assert int(f["rpp"].max()) <= depth * spp                  # the OptiX programs count rays
    if depth == 1:
        assert int(f["rpp"].sum()) == w * h * spp
        # one ray per path: the recursive programs return black at the first hit, the iterative
        # ones keep the throughput of that hit
        r2 = orc.render(orc.F32_PCG, tab, cam, w, h, spp, depth, meshes=meshes, variant=2)
        r1 = orc.render(orc.F32_PCG, tab, cam, w, h, spp, depth, meshes=meshes, variant=1)
        assert int(r1["fix"].sum() >> 32) > int(r2["fix"].sum() >> 32)


def test_eight_wide_hierarchy_on_host():
    """-DRTX_WIDTH=8 (measured and not shipped): other trees, the same frame --
    closest hits do not depend on the hierarchy."""
    sp = scenes.book1(seed=3)
    tab, meshes = scenes.table(sp, "mesh", 3)
    cam = api.camera_table(api.camera(aspratio=1.5))
    w, h, spp = 60, 40, 2
    f = orc.render(orc.F32_PCG, tab, cam, w, h, spp, 50, want_first=True, meshes=meshes)
    try:
        hostemu.use("libhostemu_w8.so")
        hostemu.stats()
        for pool in (True, False):
            e = hostemu.render(tab, cam, w, h, spp, 50, meshes=meshes, pool=pool)
            assert np.array_equal(f["first_id"], e["first_id"])
            assert np.array_equal(f["rpp"], e["rpp"])
            assert np.array_equal(f["fix"], e["fix"])
        s8 = hostemu.stats()
    finally:
        hostemu.use("libhostemu.so")
    hostemu.stats()
    for pool in (True, False):
        hostemu.render(tab, cam, w, h, spp, 50, meshes=meshes, pool=pool)
    s4 = hostemu.stats()
    assert s8["rays"] == s4["rays"] and s8["nodes"] < .8 * s4["nodes"]      # fewer, wider steps


@pytest.mark.parametrize("mode,ndiv", [("analytic", None), ("mesh", 3)])
def test_lean_pop_loop_on_host(mode, ndiv):
    """-DRTX_LEAN_POP=1 (a build option of the render kernel): the restructured
    pop loop is shared step code -- same frame, same traversal work."""
    sp = scenes.book1(seed=3)
    tab, meshes = scenes.table(sp, mode, ndiv)
    cam = api.camera_table(api.camera(aspratio=1.5))
    w, h, spp = 60, 40, 2
    f = orc.render(orc.F32_PCG, tab, cam, w, h, spp, 50, meshes=meshes)
    hostemu.stats()
    hostemu.render(tab, cam, w, h, spp, 50, meshes=meshes, want_first=False)
    s0 = hostemu.stats()
    try:
        hostemu.use("libhostemu_lp.so")
        hostemu.stats()
        e = hostemu.render(tab, cam, w, h, spp, 50, meshes=meshes, want_first=False)
        s1 = hostemu.stats()
    finally:
        hostemu.use("libhostemu.so")
    assert np.array_equal(f["rpp"], e["rpp"]) and np.array_equal(f["fix"], e["fix"])
    assert s0 == s1


def test_general_affine_instances_on_host():
    """Rotated / sheared / non-uniformly scaled mesh instances (the non-diagonal transform path)."""
    sp = scenes.affine_mix()
    tab, meshes = scenes.table(sp, "mesh")
    cam = api.camera_table(api.camera(eye=(9., 3., 6.), aspratio=1.5, aperture=.05, fostance=9.))
    w, h, spp = 60, 40, 2
    f = orc.render(orc.F32_PCG, tab, cam, w, h, spp, 50, want_first=True, meshes=meshes)
    e = hostemu.render(tab, cam, w, h, spp, 50, meshes=meshes)
    assert (f["first_id"] >= 0).mean() > .3
    assert np.array_equal(f["first_id"], e["first_id"])
    assert np.array_equal(f["rpp"], e["rpp"])
    assert np.array_equal(f["fix"], e["fix"])


def test_lbvh_traversal_equals_exhaustive_scan_on_host():
    sp = scenes.book1(seed=5)
    tab, meshes = scenes.table(sp, "mesh", 3)
    cam = api.camera_table(api.camera(aspratio=1.5))
    a = hostemu.render(tab, cam, 48, 32, 1, 0, meshes=meshes, brute=False)
    b = hostemu.render(tab, cam, 48, 32, 1, 0, meshes=meshes, brute=True)
    assert np.array_equal(a["first_id"], b["first_id"])
    assert np.array_equal(a["first_t"], b["first_t"])


def test_mesh_mode_float_mirror_tracks_double():
    sp = scenes.book1(seed=1)
    tab, meshes = scenes.table(sp, "mesh", 2)
    cam = api.camera_table(api.camera(aspratio=1.5))
    w, h, spp = 48, 32, 8
    d = orc.render(orc.F64_PCG, tab, cam, w, h, spp, 50, want_first=True, meshes=meshes)
    f = orc.render(orc.F32_PCG, tab, cam, w, h, spp, 50, want_first=True, meshes=meshes)
    assert (d["first_id"] != f["first_id"]).mean() < 2e-3
    delta = np.abs(np.clip(d["sum"] / spp, 0, 1) - orc.resolve_fix(f["fix"], spp))
    assert delta.mean() < 2e-3


def _far_small_spheres(n=60, seed=5):
    """Small spheres 100-1500 radii away from the origin region the rays start in."""
    rng = np.random.default_rng(seed)
    sp = []
    for k in range(n):
        r = float(rng.uniform(.01, .05))
        dist = r * float(rng.uniform(100., 1500.))
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        sp.append(scenes._sphere(tuple(u * dist), r, int(rng.integers(0, 3)), (.5, .5, .5), fuzz=.1, index=1.5, ndiv=2))
    return sp


def _silhouette_rays(sp, per=40, seed=6):
    """Rays from near the origin that graze the spheres: aimed at points 0.9 .. 1.1 radii off the centre."""
    rng = np.random.default_rng(seed)
    ori, d = [], []
    for s in sp:
        c, r = np.array(s["center"]), s["radius"]
        for _ in range(per):
            o = rng.uniform(-.5, .5, 3)
            ax = c - o
            perp = np.cross(ax, rng.normal(size=3))
            perp /= np.linalg.norm(perp)
            ori.append(o)
            d.append((c + perp * r * rng.uniform(.9, 1.1) - o) * rng.uniform(.2, 2.))
    return np.array(ori, dtype=np.float32), np.array(d, dtype=np.float32)


def test_bounding_sphere_pretest_is_conservative():
    """the float pre-test in front of every thing visit must never cull a ray segment that
    touches the thing's (unpadded) bounding sphere -- checked against double arithmetic for grazing rays
    at distance/radius ratios 100 .. 1500, and for the radius-1000 ground seen from its surface."""
    sp = _far_small_spheres()
    ori, d = _silhouette_rays(sp, per=200)
    per = len(ori) // len(sp)
    bs_exact = np.repeat(np.array([list(s["center"]) + [s["radius"]] for s in sp], dtype=np.float64), per, axis=0)
    bs_pad = np.repeat(np.array([hostemu.world_bsphere(scenes.xf_of(s), [0, 0, 0, 1]) for s in sp], dtype=np.float32), per, axis=0)
    rng = np.random.default_rng(1)
    tmin = np.full(len(ori), 1e-3, dtype=np.float32)
    tbest = np.where(rng.random(len(ori)) < .5, np.inf, rng.uniform(.5, 1.5, len(ori))).astype(np.float32)
    # double-precision truth: the parameter interval in which the ray is inside the unpadded sphere
    o64, d64 = ori.astype(np.float64), d.astype(np.float64)
    f = o64 - bs_exact[:, :3]
    a = (d64 * d64).sum(1)
    b = (f * d64).sum(1)
    c = (f * f).sum(1) - bs_exact[:, 3] ** 2
    disc = b * b - a * c
    touches = disc >= 0
    sq = np.sqrt(np.maximum(disc, 0))
    t0, t1 = (-b - sq) / a, (-b + sq) / a
    touches &= (t1 >= tmin) & (t0 <= tbest)
    miss = hostemu.bsphere_miss(bs_pad, ori, d, tmin, tbest)
    assert touches.sum() > 1000 and (~touches).sum() > 1000
    assert not (miss & touches).any()
    assert (miss & ~touches).mean() > .25          # ... and it still culls
    # the ground: rays that start on the radius-1000 sphere and leave at grazing angles
    n = 4000
    ang = rng.uniform(0, 2 * np.pi, n)
    p = np.stack([rng.uniform(-15, 15, n), np.zeros(n), rng.uniform(-15, 15, n)], axis=1)
    p[:, 1] = np.sqrt(1e6 - p[:, 0] ** 2 - p[:, 2] ** 2) - 1000.
    dd = np.stack([np.cos(ang), rng.uniform(-.02, .02, n), np.sin(ang)], axis=1)
    g = hostemu.world_bsphere([1000, 0, 0, 0, 0, 1000, 0, -1000, 0, 0, 1000, 0], [0, 0, 0, 1])
    o32, d32 = p.astype(np.float32), dd.astype(np.float32)
    f = o32.astype(np.float64) - np.array([0., -1000., 0.])
    a, b, c = (d32.astype(np.float64) ** 2).sum(1), (f * d32).sum(1), (f * f).sum(1) - 1e6
    disc = b * b - a * c
    t1 = (-b + np.sqrt(np.maximum(disc, 0))) / a
    touches = (disc >= 0) & (t1 >= 1e-3)
    miss = hostemu.bsphere_miss(np.tile(g, (n, 1)), o32, d32, np.full(n, 1e-3, np.float32), np.full(n, np.inf, np.float32))
    assert touches.sum() > 500 and not (miss & touches).any()


@pytest.mark.parametrize("mode", ["analytic", "mesh"])
def test_far_small_things_equal_oracle_on_host(mode):
    """Small things far from the ray origins (distance/radius 100 .. 1500): the hierarchy with its
    pre-test returns the oracle's hits for rays that graze them."""
    sp = _far_small_spheres()
    tab, meshes = scenes.table(sp, mode)
    ori, d = _silhouette_rays(sp)
    rid, rts = orc.trace_rays_f32(tab, ori, d, meshes=meshes)
    for pool in (False, True):
        ids, ts = hostemu.trace_rays(tab, ori, d, meshes=meshes, pool=pool)
        assert np.array_equal(ids, rid) and np.array_equal(ts, rts)
    assert .2 < (rid >= 0).mean() < .8
