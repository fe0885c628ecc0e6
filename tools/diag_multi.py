import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rtxplay_b200 import api, scenes
sp = scenes.book1(seed=1)
w, h = 72, 48
cam = api.camera(aspratio=w / h)
one = api.Context(0); scenes.load(one, sp, "analytic"); one.resize(w, h)
many = api.Context(devices=[0, 0]); scenes.load(many, sp, "analytic"); many.resize(w, h)
def acc(ctx, **kw):
    ctx.render_accumulate(ctx.params(cam, **kw)); return ctx.read(api.BUF_ACCUM)
a01 = acc(one, spp=2)
a0 = acc(one, spp=1, sample0=0, sample_stride=2)
a1 = acc(one, spp=1, sample0=1, sample_stride=2)
print("one: 0+1 == both", np.array_equal(a0 + a1, a01))
m = acc(many, spp=2)
print("many == one", np.array_equal(m, a01), "many == a0", np.array_equal(m, a0), "many == a1", np.array_equal(m, a1), "many == 2*a0", np.array_equal(m, 2 * a0), "many==2*a1", np.array_equal(m, 2*a1))
print("sums", m[..., 3].sum(), a01[..., 3].sum(), a0[..., 3].sum(), a1[..., 3].sum())
print(many.frame_stats())
