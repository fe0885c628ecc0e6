"""Development harness: the CUDA core (rtx_core.cuh / rtx_lbvh.cuh helpers) compiled for
the host and driven serially.  Test infrastructure only -- see hostemu.cu."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-C", _HERE, "libhostemu.so"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _LIB = ctypes.CDLL(os.path.join(_HERE, "libhostemu.so"))
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def render(things, cam, w, h, spp, depth=50, seed=4711, sample0=0, sample_stride=1, meshes=None, brute=False):
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
    fix = np.zeros((h, w, 3), dtype=np.uint64)
    rpp = np.zeros((h, w), dtype=np.uint32)
    fid = np.full((h, w), -1, dtype=np.int64)
    ft = np.zeros((h, w), dtype=np.float32)
    L.emu_render(_p(things), ctypes.c_int(len(things)), ctypes.c_int(n), vp, _p(nv), ip, _p(nt), _p(cam),
                 ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(spp), ctypes.c_int(depth), ctypes.c_uint64(seed),
                 ctypes.c_int(sample0), ctypes.c_int(sample_stride), _p(fix), _p(rpp), _p(fid), _p(ft),
                 ctypes.c_int(1 if brute else 0))
    return dict(fix=fix, rpp=rpp, first_id=fid, first_t=ft)
