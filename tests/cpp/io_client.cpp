// The match consumers and the timer printers of the public headers, exercised the same
// way against BOTH libraries: tests/golden/make_io_golden.py compiles this file against the reference's own headers
// and sources (cUtil.cc:938-954,1002-1016; Util/common.cpp:5-36) to produce the committed expectations, the test
// suite compiles it against include/3dsift + libsift3d_b200.so.  No GPU is involved.
#include "Include/cSIFT3D.h"
#include "Include/cUtil.h"

#include <cstdio>
#include <iostream>
#include <vector>

using namespace std;
using namespace CPUSIFT;

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const string out = argv[1];
    vector<Cvec> pts;
    pts.push_back(Cvec(0.0f, 1.0f, 2.0f));
    pts.push_back(Cvec(12.125f, -3.5f, 511.0f));
    pts.push_back(Cvec(1.0f / 3.0f, 2.0f / 3.0f, 100000.75f));
    pts.push_back(Cvec(1e-7f, 123456.789f, -0.000015f));
    pts.push_back(Cvec(255.99999f, 3.14159274f, 16777216.0f));
    write_sift_kp(pts, out.c_str());
    vector<Cvec> back;
    back.push_back(Cvec(9.0f, 9.0f, 9.0f));  // read_sift_kp APPENDS
    read_sift_kp(out.c_str(), back);
    cout << "READ " << back.size() << endl;
    for (size_t i = 0; i < back.size(); ++i) printf("%a %a %a\n", back[i].x, back[i].y, back[i].z);
    // an empty list: the file exists and is empty
    vector<Cvec> none;
    const string out2 = out + ".empty";
    write_sift_kp(none, out2.c_str());
    read_sift_kp(out2.c_str(), none);
    cout << "EMPTY " << none.size() << endl;

    SIFT_TimerPara t;
    t.d_TotalTime = 12.5; t.d_Allocation = 0.125; t.d_BuildGSS = 5.925; t.d_BuildDOG = 0; t.d_Detect = 0.85;
    t.d_AssignOrientation = 1.671; t.d_Extraction = 9.553; t.d_release = 0.003; t.d_memoryOverhead = 1e-5;
    cout << t;
    t.vD_octaveTime.push_back(1.5); t.vD_octaveTime.push_back(0.25);
    t.vD_octaveCompute.push_back(1.25); t.vD_octaveCompute.push_back(0.125);
    cout << t;
    cout << "COMPUTE " << t.getAllComputeTime() << endl;
    SIFT_PROCESS p;
    p.REF = t; p.TAR.d_TotalTime = 3; p.d_RegTime = 0.32;
    cout << p << endl;
    return 0;
}
