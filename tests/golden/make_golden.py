#!/usr/bin/env python
"""Regenerates tests/golden/rtow_pin.npz.  Run in the build container (needs
/root/reference): it executes the UNMODIFIED reference program oracle/_ref/rtow
(compiled from /root/reference/rtow.cxx by oracle/Makefile), records its PPM, and the
libc-rand() call counts the oracle needs to replay selected row bands in isolation.

Stored:
  md5            md5 of the reference PPM (P3 1280x720, 10 spp, depth 50)
  row_crc        crc32 of every image row of the reference PPM (row 0 = top line)
  bands          [n,2] PPM line ranges [a,b) kept verbatim in band_rgb (from the reference)
  band_skip      libc rand() calls consumed before each band's first line is rendered
  census         (diffuse, reflect, refract) thing counts of the reference scene
  stat_300x200   (paths, segments) of the 300x200x10spp run; SURVEY.md appendix A measured
                 1 625 532 segments with an instrumented copy of the reference
"""
import hashlib
import os
import subprocess
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402

W, H, SPP, DEPTH = 1280, 720, 10, 50
BANDS = [(0, 4), (330, 334), (500, 503)]   # PPM lines (0 = top = image row H-1)


def main():
    orc.build()
    exe = os.path.join(ROOT, "oracle", "_ref", "rtow")
    ppm = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    md5 = hashlib.md5(ppm).hexdigest()
    ref = np.array(ppm.split()[4:], dtype=np.int64).reshape(H, W, 3).astype(np.uint8)
    row_crc = np.array([zlib.crc32(ref[r].tobytes()) for r in range(H)], dtype=np.uint64)

    things = orc.rtow_scene()
    cam = orc.camera_f64((13, 2, 3), (0, 0, 0), (0, 1, 0), 20., 16. / 9., .1, 10.)
    cuts = sorted({0, H} | {a for a, _ in BANDS} | {b for _, b in BANDS})
    calls_at = {}
    mine = np.zeros_like(ref)
    for a, b in zip(cuts[:-1], cuts[1:]):
        calls_at[a] = orc.libc_calls()
        out = orc.render(orc.F64_LIBC, things, cam, W, H, SPP, DEPTH, y0=H - b, y1=H - a)
        mine[a:b] = orc.ppm_rtow(out["sum"], SPP)[::-1][a:b]
    assert np.array_equal(mine, ref), "oracle does not reproduce the reference image"

    t = things[:, orc.TH_TYPE]
    cam32 = orc.camera_f64((13, 2, 3), (0, 0, 0), (0, 1, 0), 20., 3. / 2., .1, 10.)
    things2 = orc.rtow_scene()
    st = orc.render(orc.F64_LIBC, things2, cam32, 300, 200, 10, 50)
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "rtow_pin.npz"),
        md5=np.array(md5), row_crc=row_crc, bands=np.array(BANDS, dtype=np.int64),
        band_skip=np.array([calls_at[a] for a, _ in BANDS], dtype=np.uint64),
        band_rgb=np.concatenate([ref[a:b] for a, b in BANDS], axis=0),
        census=np.array([(t == 0).sum(), (t == 1).sum(), (t == 2).sum()], dtype=np.int64),
        scene=things,
        stat_300x200=np.array([300 * 200 * 10, int(st["rpp"].sum())], dtype=np.int64))
    print("md5", md5, "bands", BANDS, "skip", [calls_at[a] for a, _ in BANDS],
          "segments@300x200", int(st["rpp"].sum()))


if __name__ == "__main__":
    main()
