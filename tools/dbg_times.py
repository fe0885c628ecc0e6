"""Development instrument: per-warp start / ran-dry / end times of k_render (library built with
-DRTX_DEBUG_TIMES=1, which stores them in the hit-id buffer).  usage: RTX_LIB=.../librtx_dbg.so python tools/dbg_times.py"""
import sys, numpy as np
sys.path.insert(0, '.')
from rtxplay_b200 import api, scenes
sp = scenes.book1(seed=1)
ctx = api.Context(0)
scenes.load(ctx, sp, "mesh", None)
w, h = 1200, 800
ctx.resize(w, h)
cam = api.camera(aspratio=w / h)
for spp, depth in ((16, 50), (63, 50), (63, 8), (63, 3)):
    ctx.render(ctx.params(cam, spp, depth))
    ms = ctx.last_render_ms()
    t = ctx.read(api.BUF_HIT_ID).reshape(-1)[:3 * 2960].reshape(-1, 3).astype(np.int64)
    t0 = t[:, 0].min()
    st, ex, en = (t[:, 0] - t0) / 1e6, (t[:, 1] - t0) / 1e6, (t[:, 2] - t0) / 1e6
    pc = lambda a: [round(float(np.percentile(a, q)), 2) for q in (0, 10, 50, 90, 99, 100)]
    print("spp", spp, "depth", depth, "kernel ms %.2f" % ms, "ran dry", pc(ex), "end", pc(en), "residual per warp", pc(en - ex), "segments", ctx.stats()["segments"])
ctx.close()
