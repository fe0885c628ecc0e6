"""LBVH build at scale (BASELINE configs[4]): ONE mesh of n_spheres tessellated spheres
(subdivision 6, 16 384 triangles each) flattened into world space -- a slice of the stress scene as a
single >= 100 M-primitive build -- with the device time of every build stage against the HBM
streaming bound of its traffic.  usage: build_scale.py <n_spheres> [--render]"""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from rtxplay_b200 import api  # noqa: E402

n_spheres = int(sys.argv[1])
v0, i0 = api.sphere_mesh(1., 6)
side = int(math.ceil(math.sqrt(n_spheres)))
rng = np.random.default_rng(3)
t0 = time.perf_counter()
nv, nt = len(v0), len(i0)
xyz = np.empty((n_spheres, nv, 3), dtype=np.float32)
idx = np.empty((n_spheres, nt, 3), dtype=np.uint32)
for k in range(n_spheres):
    a, b = divmod(k, side)
    c = np.array([a * .5 + .25 * rng.random(), .1, b * .5 + .25 * rng.random()], dtype=np.float32)
    xyz[k] = v0 * np.float32(.1) + c
    idx[k] = i0 + np.uint32(k * nv)
xyz = xyz.reshape(-1, 3)
idx = idx.reshape(-1, 3)
gen_s = time.perf_counter() - t0
ctx = api.Context(0)
t0 = time.perf_counter()
m = ctx.add_mesh(xyz, idx)
wall_s = time.perf_counter() - t0
st = ctx.stats()
blas, _ = ctx.build_stages()
n = len(idx)
peak = 6545.3
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]
# bytes every stage has to move per primitive (reads + writes of its arrays, each touched once per pass)
model = {"keys": 32 * 2 + 12,                 # primitive boxes read by the bounds reduction and by the key kernel; key + index written
         "sort": 8 + 8 * (12 + 12),           # keys read once for the eight histograms; 8 passes: keys + indices read and written
         "hierarchy": 8 + 16 + 16 + 8,        # sorted keys read; children, ranges, parents written
         "boxes": 4 + 32 + 64 + 64 + 4,       # order, primitive box gathered; leaf + inner boxes written (and read once by the parent); flags
         "wide_nodes": 16 + 64 + 128 / 1.5}   # binary children / ranges / boxes read; one 128-byte node per ~1.5 primitives written
out = {"primitives": n, "vertices": len(xyz), "host_generation_s": gen_s, "rtx_mesh_create_wall_s": wall_s,
       "device_ms_total": st["ms_build_blas"], "device_mb": st["bytes_device"] / 1048576., "hbm_peak_gbs": peak, "stages": {}}
for k, ms in blas.items():
    byts = model[k] * n
    out["stages"][k] = {"ms": ms, "model_bytes_per_primitive": model[k], "hbm_bound_ms": byts / (peak * 1e9) * 1e3,
                        "frac_of_hbm_bound": (byts / (peak * 1e9) * 1e3) / ms if ms > 0 else None}
out["stages_ms_sum"] = sum(blas.values())
if "--render" in sys.argv:
    # the mesh as one thing, a short frame (checks the tree end to end: every primary ray finds a surface or the sky)
    ctx.add_thing(m, api.Optics(api.DIFFUSE, (.5, .5, .5)))
    ctx.build()
    w, h = 600, 400
    ctx.resize(w, h)
    span = side * .5
    cam = api.camera(eye=(span * .5, span * .35, -span * .35), pat=(span * .5, 0., span * .5), aspratio=w / h, aperture=0., fostance=span)
    ctx.render(ctx.params(cam, 4))
    s2 = ctx.stats()
    out["render"] = {"ms": s2["ms_render"], "segments": s2["segments"], "hit_fraction": float((ctx.primary_hits(ctx.params(cam, 1))[0] >= 0).mean())}
print("BUILD " + json.dumps(out))
ctx.close()
