"""Small workload for compute-sanitizer (tools/gpu_sanitize.sh): every device entry point of
include/rtx.h once, on a tiny image, analytic and mesh scenes, all three path variants."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from rtxplay_b200 import api, scenes  # noqa: E402

spheres = scenes.book1(seed=1)
w, h = 64, 48
# (RTX_KERNEL=q in the environment runs the same through the compacting-pool kernel; the second pass
#  of each scene goes through a two-replica context: fan-out, sample split, fused reduce + resolve)
for mode, ndiv, devices in (("analytic", None, None), ("mesh", 2, None), ("mesh", 2, [0, 0])):
    ctx = api.Context(0) if devices is None else api.Context(devices=devices)
    scenes.load(ctx, spheres, mode, ndiv)
    ctx.resize(w, h)
    cam = api.camera(aspratio=w / h)
    for variant in (api.VARIANT_RTOW, api.VARIANT_RTWO_I, api.VARIANT_RTWO_R):
        p = ctx.params(cam, 3, guides=1, variant=variant)
        ctx.render(p)
    ctx.render_accumulate(ctx.params(cam, 2, sample0=3, accumulate=1))
    ctx.resolve(5)
    ctx.postproc(api.PP_SRGB)
    ctx.postproc(api.PP_NONE)
    img = ctx.read(api.BUF_IMAGE)
    ids, ts = ctx.primary_hits(p)
    ctx.pick(p, w // 2, h // 3)
    xf = ctx.get_xf(5)
    xf[3] += .25
    ctx.set_xf(5, xf)
    ctx.update()
    ctx.render(ctx.params(cam, 1))
    rng = np.random.default_rng(3)
    o = np.tile(np.array([13., 2., 3.], dtype=np.float32), (500, 1))
    d = (rng.normal(size=(500, 3)) * .2 + np.array([-13., -2., -3.])).astype(np.float32)
    a = ctx.trace_rays(o, d)
    b = ctx.trace_rays(o, d, brute=True)
    assert np.array_equal(a[0], b[0])
    ctx.probe_read(1 << 20, 2)
    print(mode, "ok", ctx.stats()["launches"], "launches", int(img.sum()))
    ctx.close()
