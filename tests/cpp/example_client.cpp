// A client shaped like the reference's 3DSIFT/Example.cpp:8-64, compiled against the drop-in
// headers and linked with libsift3d_b200.so.  Volumes come from raw "m n p + float32" files (the
// format of CreateCSIFT3D(std::string), Src/cSIFT3D.cc:112-125) or, with --mem, are loaded by the
// client and passed through the float* overload exactly as Example.cpp:21 does.
#include "Include/cSIFT3D.h"
#include "Include/cMatcher.h"

#include <cstdio>
#include <cstring>
#include <vector>

using namespace std;

static float* load(const char* path, int& nx, int& ny, int& nz) {
    FILE* f = fopen(path, "rb");
    if (!f) return nullptr;
    int d[3];
    if (fread(d, sizeof(int), 3, f) != 3) { fclose(f); return nullptr; }
    nx = d[0]; ny = d[1]; nz = d[2];
    float* v = new float[(size_t)nx * ny * nz];
    size_t got = fread(v, sizeof(float), (size_t)nx * ny * nz, f);
    fclose(f);
    if (got != (size_t)nx * ny * nz) { delete[] v; return nullptr; }
    return v;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s ref.bin tar.bin [--mem]\n", argv[0]); return 2; }
    const bool mem = argc > 3 && !strcmp(argv[3], "--mem");
    CPUSIFT::CSIFT3D *SIFT_ref, *SIFT_tar;
    float *refVol = nullptr, *tarVol = nullptr;
    if (mem) {
        int nx = 0, ny = 0, nz = 0;
        refVol = load(argv[1], nx, ny, nz);
        cout << "Dimensions of reference image:" << nx << " " << ny << " " << nz << endl;
        SIFT_ref = CPUSIFT::CSIFT3DFactory::CreateCSIFT3D(refVol, nx, ny, nz);
        int nxT = 0, nyT = 0, nzT = 0;
        tarVol = load(argv[2], nxT, nyT, nzT);
        cout << "Dimensions of target image:" << nxT << " " << nyT << " " << nzT << endl;
        SIFT_tar = CPUSIFT::CSIFT3DFactory::CreateCSIFT3D(tarVol, nxT, nyT, nzT);
    } else {
        SIFT_ref = CPUSIFT::CSIFT3DFactory::CreateCSIFT3D(std::string(argv[1]));
        SIFT_tar = CPUSIFT::CSIFT3DFactory::CreateCSIFT3D(std::string(argv[2]));
    }
    SIFT_ref->KpSiftAlgorithm();
    auto vRefKp = SIFT_ref->GetKeypoints();
    SIFT_tar->KpSiftAlgorithm();
    auto vTarKp = SIFT_tar->GetKeypoints();

    CPUSIFT::muBruteMatcher matcher;
    vector<CPUSIFT::Cvec> matchRefCoor, matchTarCoor;
    const float threshold = 0.85f;
    matcher.enhancedMatch(matchRefCoor, matchTarCoor, vRefKp, vTarKp, threshold);

    cout << "KEYPOINTS " << vRefKp.size() << " " << vTarKp.size() << endl;
    cout << "Matched Points: reference coordinate(x,y,z);target coordinate(x,y,z)" << endl;
    for (size_t i = 0; i < matchRefCoor.size(); ++i)
        cout << matchRefCoor[i].x << "," << matchRefCoor[i].y << "," << matchRefCoor[i].z << ";" << matchTarCoor[i].x << ","
             << matchTarCoor[i].y << "," << matchTarCoor[i].z << endl;
    cout << "STATUS " << SIFT_ref->LastStatus() << " " << SIFT_tar->LastStatus() << " " << matcher.LastStatus() << endl;

    delete[] refVol;
    delete[] tarVol;
    delete SIFT_ref;
    delete SIFT_tar;
    return 0;
}
