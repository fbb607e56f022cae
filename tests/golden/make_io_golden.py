"""Expectations for tests/test_io_cpu.py, produced by the REFERENCE's own code where /root/reference exists:

    python tests/golden/make_io_golden.py

tests/cpp/io_client.cpp is compiled against the reference's headers and sources (write_sift_kp / read_sift_kp,
Src/cUtil.cc:938-954,1002-1016; operator<< of SIFT_TimerPara / SIFT_PROCESS, Src/Util/common.cpp:5-36) with the oracle's
MSVC shim; its stdout and the CSV it writes are committed under tests/golden/io/."""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/3DSIFT"
OUT = os.path.join(HERE, "io")


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    os.makedirs(OUT, exist_ok=True)
    shim = os.path.join(ROOT, "oracle", "shim")
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "io_ref")
        # "Include/..." resolves inside the reference tree; only the three TUs the client needs are compiled
        subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1", "-w", "-fopenmp", "-fpermissive", "-Wno-narrowing", "-D__declspec(x)=",
                               "-include", os.path.join(shim, "msvc_shim.h"), "-I", os.path.join(shim, "x"), "-I", REF,
                               "-I", os.path.join(REF, "Include"), "-I", os.path.join(REF, "3party", "Eigen"),
                               os.path.join(ROOT, "tests", "cpp", "io_client.cpp"),
                               os.path.join(REF, "Src", "cUtil.cc"), os.path.join(REF, "Src", "Util", "common.cpp"),
                               os.path.join(REF, "Src", "Util", "cTexImage.cc"), os.path.join(REF, "Src", "cSIFT3D.cc"),
                               os.path.join(REF, "Src", "Util", "matrixIO3D.cpp"), "-o", exe])
        csv = os.path.join(td, "kp.csv")
        out = subprocess.run([exe, csv], capture_output=True, text=True, check=True).stdout
        out = out.replace(td, "<TMP>")
        open(os.path.join(OUT, "io_client.stdout"), "w").write(out)
        open(os.path.join(OUT, "kp.csv"), "w").write(open(csv).read())
    print(out)


if __name__ == "__main__":
    main()
