"""The host side of the drop-in against the reference's own tools.  tests/golden/tools.npz
holds the byte output of optx/sphere.cxx -DMAIN, optx/args.cxx -DMAIN and optx/reduce.cxx,
compiled unmodified from /root/reference (tests/golden/make_tool_golden.py)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from rtxplay_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "rtxplay_b200", "host")
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "tools.npz"))


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "rtxplay_b200", "csrc")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HOST], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def run(exe, args):
    r = subprocess.run([os.path.join(HOST, exe)] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return r.returncode, r.stdout, r.stderr


@pytest.mark.parametrize("name,args", [("sphere_1_0", ["1.", "0"]), ("sphere_1_1", ["1.", "1"]), ("sphere_1_2", ["1.", "2"]),
                                       ("sphere_1_3", ["1.", "3"]), ("sphere_02_2", [".2", "2"])])
def test_sphere_tool_writes_the_reference_scene_file(name, args):
    """Vertices, their first-appearance order, faces and the %f formatting, byte for byte."""
    assert run("sphere", args)[1] == GOLD[name].tobytes()


@pytest.mark.parametrize("ndiv", [6, 8])
def test_sphere_tool_large_meshes_by_md5(ndiv):
    """sphere_6.scn / sphere_8.scn of the reference's default scene (optx/Makefile:204-212)."""
    assert hashlib.md5(run("sphere", ["1.", str(ndiv)])[1]).hexdigest() == str(GOLD["sphere_1_%d_md5" % ndiv])


@pytest.mark.parametrize("case", range(int(GOLD["args_n"])))
def test_args_parser_reports_like_the_reference(case):
    cmd = str(GOLD["args_%d_cmd" % case])
    argv = cmd.split("\x00") if cmd else []
    rc, out, _ = run("args", argv)
    assert rc == int(GOLD["args_%d_rc" % case])
    assert out == GOLD["args_%d_out" % case].tobytes()


def test_args_rejects_unknown_options():
    rc, _, err = run("args", ["--no-such-option"])
    assert rc == 1 and b"try 'rtwo --help'" in err


def test_object_reader_round_trips_scene_files(tmp_path):
    """Object (in-house OBJ reader) on a file written by the sphere tool gives the file back
    (the reference's object.cxx -DMAIN prints the same listing)."""
    scn = tmp_path / "sphere_3.scn"
    scn.write_bytes(run("sphere", ["1.", "3"])[1])
    rc, out, _ = run("object", [str(scn)])
    assert rc == 0
    v = [l for l in out.split(b"\n") if l.startswith(b"v ")]
    f = [l for l in out.split(b"\n") if l.startswith(b"f ")]
    ref = GOLD["sphere_1_3"].tobytes().split(b"\n")
    assert v == [l for l in ref if l.startswith(b"v ")]
    assert f == [l for l in ref if l.startswith(b"f ")]


def test_dedup_matches_reduce_demo():
    """optx/reduce.cxx: 12 soup vertices -> 6 unique in first-appearance order, faces
    {0,1,2}{1,3,4}{2,4,5}{2,4,1}.  The same rule drives rtx_sphere_mesh: the first face of the
    once-subdivided tetrahedron has exactly that structure."""
    txt = GOLD["reduce_out"].tobytes().decode()
    assert "{ 0, 1, 2 }" in txt.replace("  ", " ") or "0, 1, 2" in txt
    v, i = api.sphere_mesh(1., 1)
    assert i[:4].tolist() == [[0, 1, 2], [1, 3, 4], [2, 4, 5], [1, 4, 2]]
    assert len(np.unique(v, axis=0)) == len(v)          # no duplicates survive


def test_rtwo_without_a_device_fails_like_the_reference():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    rc, out, err = run("rtwo", ["-g", "64x48", "-s", "1", "-q"])
    assert rc == 1 and err.startswith(b"exception: ")
