"""Multi-GPU through the PRODUCT path (rtx_init_multi, include/rtx.h) on real peers: the N-device
frame equals the one-device frame bit for bit, and frame times of the bench configuration (C3) and,
with --c4, of BASELINE configs[3] (3840x2160, 4096 spp).  usage: multi_check.py <n_devices> [--c4]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from rtxplay_b200 import api, scenes  # noqa: E402

n = int(sys.argv[1])
sp = scenes.book1(seed=1)
out = {"n_devices": n}
one = api.Context(0)
scenes.load(one, sp, "mesh")
t0 = time.perf_counter()
many = api.Context(devices=list(range(n)))
scenes.load(many, sp, "mesh")
out["scene_replication_s"] = time.perf_counter() - t0
w, h = 1200, 800
cam = api.camera(aspratio=w / h)
for c in (one, many):
    c.resize(w, h)
    c.render(c.params(cam, 19, guides=1))
out["bit_identical_19spp"] = bool(np.array_equal(one.read(api.BUF_ACCUM), many.read(api.BUF_ACCUM)) and np.array_equal(one.read(api.BUF_RAWRGB), many.read(api.BUF_RAWRGB))
                                  and np.array_equal(one.read(api.BUF_GUIDE_ACC), many.read(api.BUF_GUIDE_ACC)))
ms = {"one": [], "many": []}
for k in range(4):
    for name, c in (("one", one), ("many", many)):
        c.render(c.params(cam, 500))
        fs = c.frame_stats()
        if k:
            ms[name].append((fs["ms_frame"], fs["ms_trace"], fs["ms_reduce_resolve"]))
out["c3_ms_frame_1gpu"] = float(np.mean([m[0] for m in ms["one"]]))
out["c3_ms_frame"] = float(np.mean([m[0] for m in ms["many"]]))
out["c3_ms_root_kernel"] = float(np.mean([m[1] for m in ms["many"]]))
out["c3_ms_reduce_resolve"] = float(np.mean([m[2] for m in ms["many"]]))
out["c3_speedup"] = out["c3_ms_frame_1gpu"] / out["c3_ms_frame"]
out["c3_segments"] = many.stats()["segments"]
if "--c4" in sys.argv:
    one.close()
    w, h = 3840, 2160
    cam = api.camera(aspratio=w / h)
    many.resize(w, h)
    fr = []
    for k in range(2):
        many.render(many.params(cam, 4096))
        fr.append(many.frame_stats()["ms_frame"])
    st = many.stats()
    out["c4_ms_frame"] = fr
    out["c4_segments"] = st["segments"]
    out["c4_gseg_per_s"] = st["segments"] / (min(fr) * 1e-3) / 1e9
print("MULTI " + json.dumps(out))
