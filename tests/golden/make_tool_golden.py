#!/usr/bin/env python
"""Regenerates tests/golden/tools.npz from the reference's own host tools, compiled
unmodified by oracle/Makefile into oracle/_ref/ (needs /root/reference):
  sphere  optx/sphere.cxx -DMAIN : scene files (OBJ) of the subdivided tetrahedron
  args    optx/args.cxx -DMAIN   : the option parser's report for a set of command lines
  reduce  optx/reduce.cxx        : the vertex de-duplication known answer
"""
import hashlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")

ARGS_CASES = [
    ["-g", "1200x800", "-s", "500", "-d", "50", "-S"],
    ["-g", "UHD-1", "-q"],
    ["-g", "1024x0"],
    ["-g", "1200x1", "-a", "3:2"],
    ["--geometry", "4K", "--samples-per-pixel", "7", "--trace-depth", "3", "--print-aov", "RPP", "--print-guides", "-D", "NAA"],
    ["-A", "RPP,XYZ", "-v", "-t"],
    ["-s", "-12"],
    [],
]


def run(exe, args):
    r = subprocess.run([os.path.join(REF, exe)] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return r.returncode, r.stdout, r.stderr


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    out = {}
    for n in (0, 1, 2, 3):
        out["sphere_1_%d" % n] = np.frombuffer(run("sphere", ["1.", str(n)])[1], dtype=np.uint8)
    out["sphere_02_2"] = np.frombuffer(run("sphere", [".2", "2"])[1], dtype=np.uint8)
    for n in (6, 8):
        out["sphere_1_%d_md5" % n] = np.array(hashlib.md5(run("sphere", ["1.", str(n)])[1]).hexdigest())
    for k, a in enumerate(ARGS_CASES):
        rc, so, se = run("args", a)
        out["args_%d_cmd" % k] = np.array("\x00".join(a))
        out["args_%d_rc" % k] = np.array(rc)
        out["args_%d_out" % k] = np.frombuffer(so, dtype=np.uint8)
    out["args_n"] = np.array(len(ARGS_CASES))
    out["reduce_out"] = np.frombuffer(run("reduce", [])[1], dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tools.npz"), **out)
    print("wrote tools.npz:", sorted(out))


if __name__ == "__main__":
    main()
