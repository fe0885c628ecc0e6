"""Development harness: the CUDA core (rtx_core.cuh / rtx_lbvh.cuh helpers) compiled for
the host and driven serially.  Test infrastructure only -- see hostemu.cu."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _make(name):
    """(Re)build a harness library.  When the rebuild fails (a GPU-box snapshot without a writable
    tree) an existing library is used only if it is not older than the sources it is made from --
    a stale binary would let the CPU parity tests pass against code that is no longer there."""
    path = os.path.join(_HERE, name)
    r = subprocess.run(["make", "-C", _HERE, name], capture_output=True, text=True)
    if r.returncode != 0:
        if not os.path.exists(path):
            raise RuntimeError("cannot build %s:\n%s" % (name, r.stderr[-2000:]))
        src_dir = os.path.join(_HERE, "..", "..", "rtxplay_b200", "csrc")
        srcs = [os.path.join(_HERE, "hostemu.cu")] + [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cuh", ".h"))]
        if os.path.getmtime(path) < max(os.path.getmtime(f) for f in srcs):
            raise RuntimeError("make %s failed and the existing library is older than its sources:\n%s" % (name, r.stderr[-2000:]))
        import warnings
        warnings.warn("make %s failed, using the existing (up-to-date) library:\n%s" % (name, r.stderr[-500:]))
    return path


def use(name="libhostemu.so"):
    """Select a build of the harness (libhostemu_w8.so: the experimental 8-wide hierarchy)."""
    global _LIB
    _LIB = ctypes.CDLL(_make(name))
    return _LIB


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(_make("libhostemu.so"))
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def render(things, cam, w, h, spp, depth=50, seed=4711, sample0=0, sample_stride=1, meshes=None, brute=False, pool=True, want_first=True, variant=0):
    L = lib()
    L.emu_set_variant(ctypes.c_int(variant))
    things = np.ascontiguousarray(things, dtype=np.float64).reshape(-1, 20)
    cam = np.ascontiguousarray(cam, dtype=np.float64)
    meshes = meshes or []
    v = [np.ascontiguousarray(m[0], dtype=np.float32).reshape(-1, 3) for m in meshes]
    i = [np.ascontiguousarray(m[1], dtype=np.uint32).reshape(-1, 3) for m in meshes]
    n = len(meshes)
    vp = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in v])
    ip = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in i])
    nv = np.array([len(a) for a in v] or [0], dtype=np.uint32)
    nt = np.array([len(a) for a in i] or [0], dtype=np.uint32)
    fix = np.zeros((h, w, 3), dtype=np.uint64)
    rpp = np.zeros((h, w), dtype=np.uint32)
    fid = np.full((h, w), -1, dtype=np.int64)
    ft = np.zeros((h, w), dtype=np.float32)
    L.emu_render(_p(things), ctypes.c_int(len(things)), ctypes.c_int(n), vp, _p(nv), ip, _p(nt), _p(cam),
                 ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(spp), ctypes.c_int(depth), ctypes.c_uint64(seed),
                 ctypes.c_int(sample0), ctypes.c_int(sample_stride), _p(fix), _p(rpp), _p(fid if want_first else None), _p(ft if want_first else None),
                 ctypes.c_int(1 if brute else 0), ctypes.c_int(2 if pool == "q" else 1 if pool else 0))
    L.emu_set_variant(ctypes.c_int(0))
    return dict(fix=fix, rpp=rpp, first_id=fid, first_t=ft)


def stats(reset=True):
    """Traversal counters of the harness since the last reset."""
    out = np.zeros(9, dtype=np.uint64)
    lib().emu_stats(_p(out), ctypes.c_int(1 if reset else 0))
    names = ["rays", "nodes", "leaves", "tris", "things", "spheres", "enters", "pushes", "max_stack"]
    return dict(zip(names, [int(v) for v in out]))


def trace(things, cam, w, h, spp, depth=50, seed=4711, meshes=None):
    """Event trace of every path (see emu_trace in hostemu.cu): bytes."""
    L = lib()
    L.emu_trace.restype = ctypes.c_longlong
    things = np.ascontiguousarray(things, dtype=np.float64).reshape(-1, 20)
    cam = np.ascontiguousarray(cam, dtype=np.float64)
    meshes = meshes or []
    v = [np.ascontiguousarray(m[0], dtype=np.float32).reshape(-1, 3) for m in meshes]
    i = [np.ascontiguousarray(m[1], dtype=np.uint32).reshape(-1, 3) for m in meshes]
    n = len(meshes)
    vp = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in v])
    ip = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in i])
    nv = np.array([len(a) for a in v] or [0], dtype=np.uint32)
    nt = np.array([len(a) for a in i] or [0], dtype=np.uint32)
    cap = 64 << 20
    buf = ctypes.create_string_buffer(cap)
    got = L.emu_trace(_p(things), ctypes.c_int(len(things)), ctypes.c_int(n), vp, _p(nv), ip, _p(nt), _p(cam),
                      ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(spp), ctypes.c_int(depth), ctypes.c_uint64(seed),
                      buf, ctypes.c_longlong(cap))
    if got > cap:
        raise RuntimeError("trace larger than the buffer")
    return buf.raw[:got]



def warpsim(things, cam, w, h, unit_spp=64, units_per_warp=16, tile_step=97, depth=50, seed=4711, meshes=None,
            policy=0, sticky=5, t1=24, t2=33, costs=(12, 130, 40, 75, 60, 230, 150, 600, 200, 90)):
    """Warp scheduling simulator (emu_warpsim in hostemu.cu): iterations / lane-steps per step
    kind and a rough instruction estimate for the vote policy of k_render."""
    L = lib()
    things = np.ascontiguousarray(things, dtype=np.float64).reshape(-1, 20)
    cam = np.ascontiguousarray(cam, dtype=np.float64)
    meshes = meshes or []
    v = [np.ascontiguousarray(m[0], dtype=np.float32).reshape(-1, 3) for m in meshes]
    i = [np.ascontiguousarray(m[1], dtype=np.uint32).reshape(-1, 3) for m in meshes]
    n = len(meshes)
    vp = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in v])
    ip = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in i])
    nv = np.array([len(a) for a in v] or [0], dtype=np.uint32)
    nt = np.array([len(a) for a in i] or [0], dtype=np.uint32)
    out = np.zeros(24, dtype=np.float64)
    c = np.array(costs, dtype=np.float64)
    if policy == 2:
        L.emu_warpsim_pool(_p(things), ctypes.c_int(len(things)), ctypes.c_int(n), vp, _p(nv), ip, _p(nt), _p(cam),
                           ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(unit_spp), ctypes.c_int(units_per_warp), ctypes.c_int(tile_step),
                           ctypes.c_int(depth), ctypes.c_uint64(seed), ctypes.c_int(t1), ctypes.c_int(sticky), _p(c), _p(out))
    else:
      L.emu_warpsim(_p(things), ctypes.c_int(len(things)), ctypes.c_int(n), vp, _p(nv), ip, _p(nt), _p(cam),
                  ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(unit_spp), ctypes.c_int(units_per_warp), ctypes.c_int(tile_step),
                  ctypes.c_int(depth), ctypes.c_uint64(seed), ctypes.c_int(policy), ctypes.c_int(sticky), ctypes.c_int(t1), ctypes.c_int(t2),
                  _p(c), _p(out))
    names = ["", "node", "leaf", "thing", "shade", "regen", "swap", "batch"]
    r = {"rays": out[16], "paths": out[18], "cost": out[17], "cost_per_ray": out[17] / max(out[16], 1),
         "node_lines_per_ray": out[19] / max(out[16], 1)}
    for k in range(1, 8):
        if out[k]:
            r[names[k]] = (out[k] / out[16], out[8 + k] / out[k])   # iterations per ray, lanes per iteration
    return r


def _mesh_args(meshes):
    meshes = meshes or []
    v = [np.ascontiguousarray(m[0], dtype=np.float32).reshape(-1, 3) for m in meshes]
    i = [np.ascontiguousarray(m[1], dtype=np.uint32).reshape(-1, 3) for m in meshes]
    n = len(meshes)
    vp = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in v])
    ip = (ctypes.c_void_p * max(n, 1))(*[a.ctypes.data for a in i])
    nv = np.array([len(a) for a in v] or [0], dtype=np.uint32)
    nt = np.array([len(a) for a in i] or [0], dtype=np.uint32)
    return n, vp, nv, ip, nt, (v, i)


def trace_rays(things, ori, dirs, tmin=1e-3, meshes=None, brute=False, pool=False):
    """Closest hits of caller-supplied rays: the while-while traversal, the render kernel's step
    functions (pool=True; tmin must be 1e-3) or the exhaustive scan."""
    L = lib()
    things = np.ascontiguousarray(things, dtype=np.float64).reshape(-1, 20)
    ori = np.ascontiguousarray(ori, dtype=np.float32).reshape(-1, 3)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
    n, vp, nv, ip, nt, keep = _mesh_args(meshes)
    ids = np.full(len(ori), -1, dtype=np.int64)
    ts = np.zeros(len(ori), dtype=np.float32)
    L.emu_trace_rays(_p(things), ctypes.c_int(len(things)), ctypes.c_int(n), vp, _p(nv), ip, _p(nt), ctypes.c_int(len(ori)), _p(ori), _p(dirs),
                     ctypes.c_float(tmin), ctypes.c_int(1 if brute else 0), ctypes.c_int(1 if pool else 0), _p(ids), _p(ts))
    return ids, ts


def bsphere_miss(bs, ori, dirs, tmin, tbest):
    """rtx_core.cuh bsphere_miss on arrays: bs [n,4], ori/dirs [n,3], tmin/tbest [n] -> bool[n]."""
    bs, ori, dirs = (np.ascontiguousarray(a, dtype=np.float32) for a in (bs, ori, dirs))
    tmin, tbest = (np.ascontiguousarray(a, dtype=np.float32) for a in (tmin, tbest))
    out = np.zeros(len(bs), dtype=np.uint8)
    lib().emu_bsphere_miss(ctypes.c_int(len(bs)), _p(bs), _p(ori), _p(dirs), _p(tmin), _p(tbest), _p(out))
    return out.astype(bool)


def world_bsphere(xf, bs):
    xf = np.ascontiguousarray(xf, dtype=np.float32).reshape(12)
    bs = np.ascontiguousarray(bs, dtype=np.float64).reshape(4)
    out = np.zeros(4, dtype=np.float32)
    lib().emu_world_bsphere(_p(xf), _p(bs), _p(out))
    return out
