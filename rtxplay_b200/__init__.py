"""rtxplay_b200 -- B200-native path tracer behind RTXplay/RTWO's scene and launcher API.

The product is librtx.so (hand-written sm_100a CUDA kernels behind the C ABI of
include/rtx.h) plus the C++ shims in rtxplay_b200/host.  This package is the thin Python
mirror of that API used by the tests and the benchmark."""
from .api import Context, Optics, RtxError, camera, sphere_mesh  # noqa: F401
from . import scenes  # noqa: F401
