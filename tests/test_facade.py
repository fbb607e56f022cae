"""The C++ drop-in surface: an Example.cpp-shaped client compiles against include/3dsift (also
reachable as "Include/..." like the reference's own tree), links libsift3d_b200.so, and — on a GPU —
prints the same matches as the Python host mirror.  Without a GPU it must fail loudly and exit
cleanly with zero keypoints (the reference's print-and-continue convention), never fall back."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "example_client.cpp")


@pytest.fixture(scope="module")
def client(tmp_path_factory, s3d):
    exe = str(tmp_path_factory.mktemp("cpp") / "example_client")
    libdir = os.path.dirname(s3d.api.LIB_PATH)
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                           "-L", libdir, "-lsift3d_b200", f"-Wl,-rpath,{libdir}"])
    return exe


def _run(client, a, b, *extra, env=None):
    out = subprocess.run([client, a, b, *extra], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, **env) if env else None)
    return out


def test_client_compiles_and_fails_loudly_without_gpu(client, tmp_path, s3d, synth):
    if s3d.device_count() > 0:
        pytest.skip("a GPU is present")
    a, b = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    s3d.write_matrix_to_disk(a, synth.v_blobs(16, seed=1))
    s3d.write_matrix_to_disk(b, synth.v_blobs(16, seed=2))
    out = _run(client, a, b)
    assert out.returncode == 0
    assert "no CPU fallback" in out.stderr
    assert "KEYPOINTS 0 0" in out.stdout and "16 16 16" in out.stdout
    st = [l for l in out.stdout.splitlines() if l.startswith("STATUS")][0].split()
    assert st[1] != "0" and st[2] != "0"


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["file", "--mem"])
def test_client_matches_python_api(client, tmp_path, s3d, synth, mode):
    va, vb = synth.v_blobs_pair(64, seed=0)
    a, b = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    s3d.write_matrix_to_disk(a, va)
    s3d.write_matrix_to_disk(b, vb)
    out = _run(client, a, b, *([mode] if mode == "--mem" else []))
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    kps = []
    for v in (va, vb):
        s = s3d.CSIFT3DFactory.CreateCSIFT3D(v)
        s.KpSiftAlgorithm()
        kps.append((s, s.GetKeypoints()))
    m = s3d.muBruteMatcher()
    rm, tm = m.enhancedMatch(kps[0][1], kps[1][1], 0.85)
    assert f"KEYPOINTS {len(kps[0][1])} {len(kps[1][1])}" in lines
    i = lines.index("Matched Points: reference coordinate(x,y,z);target coordinate(x,y,z)")
    got = [l for l in lines[i + 1:] if ";" in l]
    want = [f"{r[0]:g},{r[1]:g},{r[2]:g};{t[0]:g},{t[1]:g},{t[2]:g}" for r, t in zip(rm, tm)]
    assert got == want and len(got) > 5
    assert "STATUS 0 0 0" in lines


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0,0", "0,0,0"])
def test_client_multi_device_env_gives_the_single_device_output(client, tmp_path, s3d, synth, devices):
    """SIFT3D_B200_DEVICES spreads CreateCSIFT3D + KpSiftAlgorithm (z-slabs, s3d_extract_multi) and the matcher
    (database shards, s3d_match_multi) of an UNCHANGED Example.cpp-shaped client over the listed devices; the same
    device listed several times = logical shards, which a one-GPU box can run.  Output must not change by a character."""
    va, vb = synth.v_blobs_pair((64, 64, 128), seed=3)
    a, b = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    s3d.write_matrix_to_disk(a, va)
    s3d.write_matrix_to_disk(b, vb)
    one = _run(client, a, b, "--mem")
    many = _run(client, a, b, "--mem", env={"SIFT3D_B200_DEVICES": devices, "S3D_SLAB_MIN_NZ": "0"})
    assert one.returncode == 0 and many.returncode == 0, many.stderr
    assert "STATUS 0 0 0" in many.stdout
    assert one.stdout == many.stdout
    assert len([l for l in many.stdout.splitlines() if ";" in l]) > 5

