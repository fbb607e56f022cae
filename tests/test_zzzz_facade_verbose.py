"""The façade's verbose mode.  Added after round 2's GPU allowance had ended: it has not run on a GPU yet,
so the file sorts last and cannot stop the rest of a `-x` run."""
import os

import pytest

from test_facade import _run, client  # noqa: F401  (the compiled Example.cpp-shaped client)


@pytest.mark.gpu
def test_verbose_mode_prints_the_reference_stage_lines(client, tmp_path, s3d, synth):
    """SIFT3D_B200_VERBOSE=1 reproduces the shape of the reference's run-time prints (time_info Src/cSIFT3D.cc:78-101 and
    the count lines of KpSiftAlgorithm :165-235) for scripts that scrape them; the default is quiet."""
    import re
    va, vb = synth.v_blobs_pair(64, seed=0)
    a, b = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    s3d.write_matrix_to_disk(a, va)
    s3d.write_matrix_to_disk(b, vb)
    quiet = _run(client, a, b, "--mem")
    loud = _run(client, a, b, "--mem", env={"SIFT3D_B200_VERBOSE": "1"})
    assert quiet.returncode == 0 and loud.returncode == 0, loud.stderr
    assert "Initialization is OK" not in quiet.stdout
    lines = loud.stdout.splitlines()
    stage = [l for l in lines if re.match(r"^\t\ttime:[-+0-9.e]+ms  --------", l)]
    names = [l.split("--------", 1)[1] for l in stage]
    assert names[:7] == ["start", "Init done", "Build GSS", "Build DOG", "Detect keypoint", "Orientation", "Description"]
    assert lines.count("Initialization is OK") == 2
    assert len([l for l in lines if re.match(r"^\ttotal time:[-+0-9.e]+s  ----finish$", l)]) == 2
    assert len([l for l in lines if l.startswith("After detecting keypoints, kp size is : ")]) == 2
    assert len([l for l in lines if l.startswith("After Orientation, kp size is : ")]) == 2
