"""ctypes binding of librtx.so (include/rtx.h).  Loading fails loudly when the library has
not been built; there is no Python or CPU fallback for any device entry point."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("RTX_LIB") or os.path.join(_HERE, "librtx.so")   # RTX_LIB: an experimental build
_LIB = None


class RtxOptics(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int32), ("albedo", ctypes.c_float * 3), ("fuzz", ctypes.c_float), ("index", ctypes.c_float)]


class RtxCamera(ctypes.Structure):
    _fields_ = [("eye", ctypes.c_float * 3), ("u", ctypes.c_float * 3), ("v", ctypes.c_float * 3),
                ("hvec", ctypes.c_float * 3), ("wvec", ctypes.c_float * 3), ("dvec", ctypes.c_float * 3),
                ("aperture", ctypes.c_float)]


class RtxParams(ctypes.Structure):
    _fields_ = [("image_w", ctypes.c_uint32), ("image_h", ctypes.c_uint32), ("spp", ctypes.c_uint32),
                ("depth", ctypes.c_uint32), ("camera", RtxCamera), ("seed", ctypes.c_uint64),
                ("sample0", ctypes.c_uint32), ("sample_stride", ctypes.c_uint32), ("accumulate", ctypes.c_uint32),
                ("guides", ctypes.c_uint32), ("variant", ctypes.c_uint32)]


class RtxStats(ctypes.Structure):
    _fields_ = [("segments", ctypes.c_uint64), ("paths", ctypes.c_uint64), ("ms_render", ctypes.c_float),
                ("ms_build_blas", ctypes.c_float), ("ms_build_tlas", ctypes.c_float), ("launches", ctypes.c_uint32),
                ("n_things", ctypes.c_uint32), ("n_meshes", ctypes.c_uint32), ("n_triangles", ctypes.c_uint64),
                ("n_triangles_instanced", ctypes.c_uint64), ("bytes_device", ctypes.c_uint64)]


class RtxFrameStats(ctypes.Structure):
    _fields_ = [("ms_frame", ctypes.c_float), ("ms_trace", ctypes.c_float), ("ms_reduce_resolve", ctypes.c_float), ("ms_postproc", ctypes.c_float),
                ("n_devices", ctypes.c_uint32), ("kernel", ctypes.c_uint32), ("counted", ctypes.c_uint32), ("pad", ctypes.c_uint32),
                ("steps", ctypes.c_uint64 * 8), ("lanes", ctypes.c_uint64 * 8), ("live_paths", ctypes.c_uint64 * 64)]


# every symbol include/rtx.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "rtx_init", "rtx_init_multi", "rtx_device_count", "rtx_shutdown", "rtx_last_error", "rtx_mesh_create", "rtx_sphere_create", "rtx_thing_add",
    "rtx_thing_set_xf", "rtx_thing_get_xf", "rtx_thing_set_optics", "rtx_accel_build", "rtx_accel_refit",
    "rtx_resize", "rtx_render", "rtx_render_accumulate", "rtx_resolve", "rtx_pick", "rtx_postproc",
    "rtx_postproc_dev", "rtx_primary_hits", "rtx_trace_rays", "rtx_read", "rtx_device_ptr", "rtx_write",
    "rtx_stats_get", "rtx_frame_stats_get", "rtx_probe_read", "rtx_build_stages", "rtx_last_render_ms", "rtx_counters_get", "rtx_camera_set", "rtx_sphere_mesh",
]


def build(force=False):
    """Compile librtx.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc")
    newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src))
    newest = max(newest, os.path.getmtime(os.path.join(_HERE, "..", "include", "rtx.h")))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        subprocess.check_call(["make", "-C", src], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            raise RuntimeError("librtx.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "or `make -C rtxplay_b200/csrc` (there is no fallback path)")
        L = ctypes.CDLL(_SO)
        L.rtx_last_error.restype = ctypes.c_char_p
        L.rtx_last_error.argtypes = [ctypes.c_void_p]
        for name in SYMBOLS:
            getattr(L, name)
        _LIB = L
    return _LIB
