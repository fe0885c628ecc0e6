"""ctypes front-end of the CPU oracle (oracle/oracle.cxx).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(rtxplay_b200) never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

F64_LIBC, F64_PCG, F32_PCG = 0, 1, 2

TH_KIND, TH_MESH, TH_XF, TH_TYPE, TH_ALB, TH_FUZZ, TH_INDEX, TH_STRIDE = 0, 1, 2, 14, 15, 18, 19, 20
CAM_STRIDE = 19


def build(force=False):
    """Compile liboracle.so (and oracle/_ref/rtow when /root/reference exists)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cxx")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/rtow.cxx") and not os.path.exists(os.path.join(_HERE, "_ref", "rtow")):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = ctypes.CDLL(so)
        L.orc_rtow_scene.restype = ctypes.c_int
        L.orc_libc_calls.restype = ctypes.c_uint64
        L.orc_libc_reset.argtypes = [ctypes.c_uint64]
        L.orc_render.restype = ctypes.c_int
        L.orc_trace_rays_f32.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a, t=ctypes.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def rtow_scene():
    """rtow.cxx:51-80 replayed with the libc stream from its unseeded start."""
    L = lib()
    L.orc_libc_reset(0)
    buf = np.zeros((600, TH_STRIDE), dtype=np.float64)
    n = L.orc_rtow_scene(_p(buf), 600)
    return buf[:n].copy()


def camera_f64(eye, pat, vup, fov, aspratio, aperture, fostance):
    cam = np.zeros(CAM_STRIDE, dtype=np.float64)
    e, p, u = (np.asarray(v, dtype=np.float64) for v in (eye, pat, vup))
    lib().orc_camera_set_f64(_p(e), _p(p), _p(u), ctypes.c_double(fov), ctypes.c_double(aspratio),
                             ctypes.c_double(aperture), ctypes.c_double(fostance), _p(cam))
    return cam


def camera_f32(eye, pat, vup, fov, aspratio, aperture, fostance):
    cam = np.zeros(CAM_STRIDE, dtype=np.float64)
    e, p, u = (np.asarray(v, dtype=np.float32) for v in (eye, pat, vup))
    lib().orc_camera_set_f32(_p(e), _p(p), _p(u), ctypes.c_float(fov), ctypes.c_float(aspratio),
                             ctypes.c_float(aperture), ctypes.c_float(fostance), _p(cam))
    return cam


class _Meshes:
    def __init__(self, meshes):
        meshes = meshes or []
        self.v = [np.ascontiguousarray(m[0], dtype=np.float32).reshape(-1, 3) for m in meshes]
        self.i = [np.ascontiguousarray(m[1], dtype=np.uint32).reshape(-1, 3) for m in meshes]
        n = len(meshes)
        self.n = n
        self.vp = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in self.v])
        self.ip = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in self.i])
        self.nv = np.array([len(a) for a in self.v] or [0], dtype=np.uint32)
        self.nt = np.array([len(a) for a in self.i] or [0], dtype=np.uint32)


def render(kind, things, cam, w, h, spp, depth=50, seed=4711, sample0=0, sample_stride=1,
           y0=0, y1=None, threads=0, meshes=None, want_first=False, want_guides=False, variant=0):
    """Returns dict(sum=double[h,w,3], fix=uint64[h,w,3], rpp=uint32[h,w], first_id, first_t).
    variant: 0 rtow.cxx semantics (the parity target), 1 / 2 the iterative / recursive OptiX
    programs where they differ (oracle.cxx Tracer::variant)."""
    L = lib()
    L.orc_set_variant(ctypes.c_int(variant))
    y1 = h if y1 is None else y1
    things = np.ascontiguousarray(things, dtype=np.float64).reshape(-1, TH_STRIDE)
    cam = np.ascontiguousarray(cam, dtype=np.float64)
    M = _Meshes(meshes)
    out = dict(sum=np.zeros((h, w, 3), dtype=np.float64), fix=np.zeros((h, w, 3), dtype=np.uint64),
               rpp=np.zeros((h, w), dtype=np.uint32))
    fid = np.full((h, w), -1, dtype=np.int64) if want_first else None
    ft = np.full((h, w), -1.0, dtype=np.float64) if want_first else None
    gd = np.zeros((h, w, 6), dtype=np.int64) if want_guides else None
    rc = L.orc_render(ctypes.c_int(kind), _p(things), ctypes.c_int(len(things)),
                      ctypes.c_int(M.n), M.vp, _p(M.nv), M.ip, _p(M.nt),
                      _p(cam), ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(spp), ctypes.c_int(depth),
                      ctypes.c_uint64(seed), ctypes.c_int(sample0), ctypes.c_int(sample_stride),
                      ctypes.c_int(y0), ctypes.c_int(y1), ctypes.c_int(threads),
                      _p(out["sum"]), _p(out["fix"]), _p(out["rpp"]), _p(fid), _p(ft), _p(gd))
    L.orc_set_variant(ctypes.c_int(0))
    if rc != 0:
        raise RuntimeError("orc_render failed: %d" % rc)
    out["first_id"], out["first_t"], out["guide"] = fid, ft, gd
    return out


def trace_rays_f32(things, ori, dirs, tmin=1e-3, threads=0, meshes=None):
    L = lib()
    things = np.ascontiguousarray(things, dtype=np.float64).reshape(-1, TH_STRIDE)
    ori = np.ascontiguousarray(ori, dtype=np.float32).reshape(-1, 3)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
    n = len(ori)
    M = _Meshes(meshes)
    ids = np.full(n, -1, dtype=np.int64)
    ts = np.zeros(n, dtype=np.float32)
    L.orc_trace_rays_f32(_p(things), ctypes.c_int(len(things)), ctypes.c_int(M.n), M.vp, _p(M.nv), M.ip, _p(M.nt),
                         ctypes.c_int(n), _p(ori), _p(dirs), ctypes.c_float(tmin), ctypes.c_int(threads), _p(ids), _p(ts))
    return ids, ts


def path_log(kind, things, cam, w, h, x, y, sample, depth=50, seed=4711, meshes=None, max_segments=128):
    """Per-segment record of one path: array [n,6] = (thing, prim, t, px, py, pz), and its colour."""
    L = lib()
    things = np.ascontiguousarray(things, dtype=np.float64).reshape(-1, TH_STRIDE)
    cam = np.ascontiguousarray(cam, dtype=np.float64)
    M = _Meshes(meshes)
    log = np.zeros((max_segments, 6), dtype=np.float64)
    rgb = np.zeros(3, dtype=np.float64)
    n = L.orc_path_log(ctypes.c_int(kind), _p(things), ctypes.c_int(len(things)), ctypes.c_int(M.n), M.vp, _p(M.nv), M.ip, _p(M.nt),
                       _p(cam), ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(depth), ctypes.c_uint64(seed),
                       ctypes.c_int(x), ctypes.c_int(y), ctypes.c_int(sample), _p(log), ctypes.c_int(max_segments), _p(rgb))
    return log[:min(n, max_segments)], rgb


def ppm_rtow(sum_, spp):
    """rtow.cxx:6-21 quantisation of per-pixel double sums -> uint8[h,w,3] (row y = image row y, y up)."""
    s = np.ascontiguousarray(sum_, dtype=np.float64)
    out = np.zeros(s.shape, dtype=np.uint8)
    lib().orc_ppm_rtow(_p(s), ctypes.c_int(spp), ctypes.c_size_t(s.size // 3), _p(out))
    return out


def resolve_fix(fix, total_spp):
    f = np.ascontiguousarray(fix, dtype=np.uint64)
    out = np.zeros(f.shape, dtype=np.float32)
    lib().orc_resolve_fix(_p(f), ctypes.c_uint64(total_spp), ctypes.c_size_t(f.size // 3), _p(out))
    return out


def srgb8(raw, srgb=True):
    r = np.ascontiguousarray(raw, dtype=np.float32)
    out = np.zeros(r.shape[:-1] + (4,), dtype=np.uint8)
    lib().orc_srgb8(_p(r), ctypes.c_size_t(r.size // 3), ctypes.c_int(1 if srgb else 0), _p(out))
    return out


def libc_reset(skip=0):
    lib().orc_libc_reset(ctypes.c_uint64(skip))


def libc_calls():
    return int(lib().orc_libc_calls())
