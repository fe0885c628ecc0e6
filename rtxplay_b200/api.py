"""Python mirror of the reference's host API over the C ABI (include/rtx.h).

Context plays the parts of Scene (optx/scene.h:19-36: add / set / get / build / update) and
Launcher (optx/launcher.h:17-24: resize / ignite) in one object, with the same argument
meaning and the same error convention: a failing call raises (the reference throws
std::runtime_error, optx/util_cpu.h:18-53)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import RtxCamera, RtxFrameStats, RtxOptics, RtxParams, RtxStats

DIFFUSE, REFLECT, REFRACT = 0, 1, 2
BUF_ACCUM, BUF_RAWRGB, BUF_RPP, BUF_IMAGE, BUF_HIT_ID, BUF_HIT_T, BUF_NORMALS, BUF_ALBEDOS, BUF_PICK_ID, BUF_GUIDE_ACC = range(10)
PP_NONE, PP_SRGB = 0, 1
VARIANT_RTOW, VARIANT_RTWO_I, VARIANT_RTWO_R = 0, 1, 2   # include/rtx.h RTX_VARIANT_*


class RtxError(RuntimeError):
    pass


def Optics(type, albedo=(0., 0., 0.), fuzz=0., index=0.):
    o = RtxOptics()
    o.type = int(type)
    o.albedo[:] = [float(a) for a in albedo]
    o.fuzz = float(fuzz)
    o.index = float(index)
    return o


def camera(eye=(13., 2., 3.), pat=(0., 0., 0.), vup=(0., 1., 0.), fov=20., aspratio=1.5, aperture=.1, fostance=10.):
    """Camera::set (optx/camera.h:30-48); defaults are rtwo's (optx/rtwo.cxx:100-108)."""
    cam = RtxCamera()
    f3 = ctypes.c_float * 3
    _lib.lib().rtx_camera_set(ctypes.byref(cam), f3(*eye), f3(*pat), f3(*vup), ctypes.c_float(fov),
                              ctypes.c_float(aspratio), ctypes.c_float(aperture), ctypes.c_float(fostance))
    return cam


def camera_table(cam):
    """The 19 numbers of a camera in the order the oracle tables use."""
    return np.array(list(cam.eye) + list(cam.u) + list(cam.v) + list(cam.hvec) + list(cam.wvec) + list(cam.dvec)
                    + [cam.aperture], dtype=np.float64)


def sphere_mesh(radius=1., ndiv=6):
    """Sphere(radius, ndiv).mesh() (optx/sphere.cxx:28-106): float32 [nv,3], uint32 [nt,3]."""
    L = _lib.lib()
    nv, nt = ctypes.c_uint32(0), ctypes.c_uint32(0)
    if L.rtx_sphere_mesh(ctypes.c_float(radius), ctypes.c_uint32(ndiv), None, ctypes.byref(nv), None, ctypes.byref(nt)):
        raise RtxError("rtx_sphere_mesh: bad subdivision count")
    xyz = np.zeros((nv.value, 3), dtype=np.float32)
    idx = np.zeros((nt.value, 3), dtype=np.uint32)
    rc = L.rtx_sphere_mesh(ctypes.c_float(radius), ctypes.c_uint32(ndiv), xyz.ctypes.data_as(ctypes.c_void_p),
                           ctypes.byref(nv), idx.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nt))
    if rc:
        raise RtxError("rtx_sphere_mesh failed (%d)" % rc)
    return xyz[:nv.value], idx


class Context:
    def __init__(self, device=0, devices=None):
        """devices: a list of CUDA device indices -> one context over all of them (rtx_init_multi):
        the scene is replicated, a frame's samples are split, the sums reduced on the first device."""
        self._L = _lib.lib()
        self._c = ctypes.c_void_p()
        if devices is None:
            rc = self._L.rtx_init(ctypes.c_int(device), ctypes.byref(self._c))
        else:
            ids = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
            rc = self._L.rtx_init_multi(ctypes.c_int(len(devices)), ids, ctypes.byref(self._c))
        if rc:
            raise RtxError(self._L.rtx_last_error(None).decode())
        self.w = self.h = 0

    def device_count(self):
        return int(self._L.rtx_device_count(self._c))

    def close(self):
        if self._c:
            self._L.rtx_shutdown(self._c)
            self._c = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise RtxError(self._L.rtx_last_error(self._c).decode())

    # ---- Scene ----
    def add_mesh(self, xyz, idx):
        """Scene::add(Object&): returns the mesh ("GAS") id."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1, 3)
        mid = ctypes.c_uint32()
        self._ck(self._L.rtx_mesh_create(self._c, xyz.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(len(xyz)),
                                         idx.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(len(idx)), ctypes.byref(mid)))
        return mid.value

    def add_analytic_sphere(self):
        mid = ctypes.c_uint32()
        self._ck(self._L.rtx_sphere_create(self._c, ctypes.byref(mid)))
        return mid.value

    def add_thing(self, mesh_id, optics, xf=None):
        """Scene::add(Thing&, object) followed by Scene::set(id, transform)."""
        tid = ctypes.c_uint32()
        self._ck(self._L.rtx_thing_add(self._c, ctypes.c_uint32(mesh_id), ctypes.byref(optics), ctypes.byref(tid)))
        if xf is not None:
            self.set_xf(tid.value, xf)
        return tid.value

    def set_xf(self, thing, xf):
        a = (ctypes.c_float * 12)(*[float(v) for v in np.asarray(xf).reshape(-1)])
        self._ck(self._L.rtx_thing_set_xf(self._c, ctypes.c_uint32(thing), a))

    def get_xf(self, thing):
        a = (ctypes.c_float * 12)()
        self._ck(self._L.rtx_thing_get_xf(self._c, ctypes.c_uint32(thing), a))
        return np.array(list(a), dtype=np.float32)

    def set_optics(self, thing, optics):
        self._ck(self._L.rtx_thing_set_optics(self._c, ctypes.c_uint32(thing), ctypes.byref(optics)))

    def build(self):
        self._ck(self._L.rtx_accel_build(self._c))

    def update(self):
        self._ck(self._L.rtx_accel_refit(self._c))

    # ---- Launcher ----
    def resize(self, w, h):
        self._ck(self._L.rtx_resize(self._c, ctypes.c_uint32(w), ctypes.c_uint32(h)))
        self.w, self.h = w, h

    def params(self, cam, spp, depth=50, seed=4711, sample0=0, sample_stride=1, accumulate=0, guides=0, variant=0):
        p = RtxParams()
        p.image_w, p.image_h, p.spp, p.depth = self.w, self.h, spp, depth
        p.camera = cam
        p.seed, p.sample0, p.sample_stride, p.accumulate = seed, sample0, sample_stride, accumulate
        p.guides, p.variant = guides, variant
        return p

    def render(self, p):
        """Launcher::ignite: blocking; rawRGB and rpp are valid afterwards."""
        self._ck(self._L.rtx_render(self._c, ctypes.byref(p)))

    def render_accumulate(self, p):
        self._ck(self._L.rtx_render_accumulate(self._c, ctypes.byref(p)))

    def resolve(self, total_spp):
        self._ck(self._L.rtx_resolve(self._c, ctypes.c_uint64(total_spp)))

    def pick(self, p, x, y):
        tid = ctypes.c_uint32()
        self._ck(self._L.rtx_pick(self._c, ctypes.byref(p), ctypes.c_uint32(x), ctypes.c_uint32(y), ctypes.byref(tid)))
        return None if tid.value == 0xffffffff else tid.value

    def postproc(self, kind=PP_SRGB):
        self._ck(self._L.rtx_postproc(self._c, ctypes.c_int(kind)))

    def primary_hits(self, p):
        self._ck(self._L.rtx_primary_hits(self._c, ctypes.byref(p)))
        return self.read(BUF_HIT_ID), self.read(BUF_HIT_T)

    def trace_rays(self, ori, dirs, tmin=1e-3, brute=False):
        ori = np.ascontiguousarray(ori, dtype=np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        ids = np.full(len(ori), -1, dtype=np.int64)
        ts = np.zeros(len(ori), dtype=np.float32)
        self._ck(self._L.rtx_trace_rays(self._c, ctypes.c_uint32(len(ori)), ori.ctypes.data_as(ctypes.c_void_p),
                                        dirs.ctypes.data_as(ctypes.c_void_p), ctypes.c_float(tmin),
                                        ctypes.c_int(1 if brute else 0), ids.ctypes.data_as(ctypes.c_void_p),
                                        ts.ctypes.data_as(ctypes.c_void_p)))
        return ids, ts

    _SHAPES = {BUF_ACCUM: (np.uint64, 4), BUF_RAWRGB: (np.float32, 3), BUF_RPP: (np.uint32, 0), BUF_IMAGE: (np.uint8, 4),
               BUF_HIT_ID: (np.int64, 0), BUF_HIT_T: (np.float32, 0), BUF_NORMALS: (np.float32, 3), BUF_ALBEDOS: (np.float32, 3),
               BUF_GUIDE_ACC: (np.int64, 6)}

    def read(self, buffer):
        dt, ch = self._SHAPES[buffer]
        shape = (self.h, self.w, ch) if ch else (self.h, self.w)
        out = np.zeros(shape, dtype=dt)
        self._ck(self._L.rtx_read(self._c, ctypes.c_int(buffer), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(out.nbytes)))
        return out

    def write(self, buffer, arr):
        dt, _ = self._SHAPES[buffer]
        a = np.ascontiguousarray(arr, dtype=dt)
        self._ck(self._L.rtx_write(self._c, ctypes.c_int(buffer), a.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(a.nbytes)))

    def device_ptr(self, buffer):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        self._ck(self._L.rtx_device_ptr(self._c, ctypes.c_int(buffer), ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def last_render_ms(self):
        ms = ctypes.c_float()
        self._ck(self._L.rtx_last_render_ms(self._c, ctypes.byref(ms)))
        return ms.value

    def probe_read(self, nbytes, repeats):
        """GB/s of all SMs reading an nbytes buffer `repeats` times past L1 (L2 peak if it fits L2)."""
        g = ctypes.c_float()
        self._ck(self._L.rtx_probe_read(self._c, ctypes.c_size_t(nbytes), ctypes.c_uint32(repeats), ctypes.byref(g)))
        return g.value

    COUNTERS = ("rays", "nodes", "leaves", "tris", "things", "culled_or_sphere_tests", "entries")

    def counters(self, reset=True):
        """Traversal events since the last reset (instrumented build librtx_count.so only)."""
        out = (ctypes.c_uint64 * 8)()
        self._ck(self._L.rtx_counters_get(self._c, out, ctypes.c_int(1 if reset else 0)))
        return dict(zip(self.COUNTERS, [int(v) for v in out]))

    BUILD_STAGES = ("keys", "sort", "hierarchy", "boxes", "wide_nodes")

    def build_stages(self):
        """Device ms per hierarchy-build stage: all meshes so far, and the last top-level build."""
        b, t = (ctypes.c_float * 5)(), (ctypes.c_float * 5)()
        self._ck(self._L.rtx_build_stages(self._c, b, t))
        return dict(zip(self.BUILD_STAGES, [float(v) for v in b])), dict(zip(self.BUILD_STAGES, [float(v) for v in t]))

    STEP_KINDS = ("", "node", "leaf", "thing", "shade", "regen")

    def frame_stats(self, reset=True):
        """Stage times of the last frame; with the instrumented build also warp iterations / lanes per
        step kind and the paths alive per bounce (rtx_frame_stats of include/rtx.h)."""
        f = RtxFrameStats()
        self._ck(self._L.rtx_frame_stats_get(self._c, ctypes.byref(f), ctypes.c_int(1 if reset else 0)))
        out = {k: getattr(f, k) for k in ("ms_frame", "ms_trace", "ms_reduce_resolve", "ms_postproc", "n_devices", "kernel", "counted")}
        if f.counted:
            out["steps"] = {self.STEP_KINDS[k]: int(f.steps[k]) for k in range(1, 6)}
            out["lanes_per_step"] = {self.STEP_KINDS[k]: (f.lanes[k] / f.steps[k] if f.steps[k] else 0.) for k in range(1, 6)}
            live = [int(v) for v in f.live_paths]
            while live and live[-1] == 0:
                live.pop()
            out["live_paths"] = live
        return out

    def stats(self):
        s = RtxStats()
        self._ck(self._L.rtx_stats_get(self._c, ctypes.byref(s)))
        return {k: getattr(s, k) for k, _ in RtxStats._fields_}
