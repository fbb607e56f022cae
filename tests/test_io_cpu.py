"""Match consumers and timer printers of the public headers: write_sift_kp / read_sift_kp
(Src/cUtil.cc:938-954,1002-1016) and operator<< of SIFT_TimerPara / SIFT_PROCESS (Src/Util/common.cpp:5-36).
tests/cpp/io_client.cpp, compiled against the drop-in headers and libsift3d_b200.so, must print and write exactly what
the same source printed and wrote when it was compiled against the reference (tests/golden/io/, produced by
tests/golden/make_io_golden.py).  Host-only functions: no GPU needed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "io")


def test_io_client_prints_and_writes_what_the_reference_does(tmp_path, s3d):
    exe = str(tmp_path / "io_client")
    libdir = os.path.dirname(s3d.api.LIB_PATH)
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "io_client.cpp"), "-o", exe, "-L", libdir, "-lsift3d_b200",
                           f"-Wl,-rpath,{libdir}"])
    csv = str(tmp_path / "kp.csv")
    out = subprocess.run([exe, csv], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    got = out.stdout.replace(str(tmp_path), "<TMP>")
    want = open(os.path.join(GOLD, "io_client.stdout")).read()
    assert got == want
    assert open(csv).read() == open(os.path.join(GOLD, "kp.csv")).read()
    assert open(csv + ".empty").read() == ""


def test_public_struct_types_are_declared(tmp_path):
    """Tri / Mesh / Image / EigenVal of Include/cSIFT3D.h:77-116 and the exported part of Include/cUtil.h compile."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "Include/cUtil.h"\n#include "Include/cMatcher.h"\n'
                   "int main() { CPUSIFT::Tri t; CPUSIFT::Mesh m; m.tri = &t; m.num = 1; CPUSIFT::Image im; im.nx = 1; CPUSIFT::EigenVal e; e.val = 0;\n"
                   " static_assert(sizeof(CPUSIFT::Tri) == 48 && sizeof(CPUSIFT::EigenVal) == 16 && sizeof(CPUSIFT::Keypoint) == 176 && sizeof(CPUSIFT::Cvec) == 12, \"layouts\");\n"
                   " return (int)(t.idx[0] * 0 + im.nx - 1 + (int)e.val + m.num - 1); }\n")
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])
