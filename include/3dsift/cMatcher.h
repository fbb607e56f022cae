// 3dsift/cMatcher.h — drop-in C++ surface of the brute-force matcher
// (reference: /root/reference/3DSIFT/Include/cMatcher.h:12-88, Src/cMatcher.cc).
#ifndef S3D_FACADE_CMATCHER_H
#define S3D_FACADE_CMATCHER_H

#include <vector>

#include "Util/common.h"
#include "cSIFT3D.h"

#define DESC_LENGTH 768

namespace CPUSIFT {

class SIFT_LIBRARY_API muBruteMatcher {
public:
    // seconds, same meaning as cMatcher.h:60-67 (device time of the corresponding kernels)
    float matchTime = 0.0f, filterTime = 0.0f, countMatchedTime = 0.0f, revMatchTime = 0.0f, revFilterTime = 0.0f,
          bijectFilterTime = 0.0f, converseTime = 0.0f, totalTime = 0.0f;

    muBruteMatcher();
    float getCalculationTime();
    std::vector<float> getGlodenDistSquare();
    std::vector<float> getSilverDistSquare();
    std::vector<int> getGlodenIdx();     // post-filter: rejected matches are negated (Src/cMatcher.cc:92-94,141-142)
    std::vector<int> getSilverIdx();

    // Matched coordinate pairs (rx, ry, rz) are APPENDED to refMatch / tarMatch in ascending
    // reference order (toCvec, Src/cMatcher.cc:99-112).
    void injectMatch(std::vector<Cvec>& refMatch, std::vector<Cvec>& tarMatch, const std::vector<Keypoint>& ref_kp,
                     const std::vector<Keypoint>& tar_kp, const double thresHold = 0.85);
    void bijectMatch(std::vector<Cvec>& refMatch, std::vector<Cvec>& tarMatch, const std::vector<Keypoint>& ref_kp,
                     const std::vector<Keypoint>& tar_kp, const double thresHold = 0.85);
    void enhancedMatch(std::vector<Cvec>& refMatch, std::vector<Cvec>& tarMatch, const std::vector<Keypoint>& ref_kp,
                       const std::vector<Keypoint>& tar_kp, const double thresHold = 0.85);

    int LastStatus() const { return status; }

private:
    void run(int type, std::vector<Cvec>& refMatch, std::vector<Cvec>& tarMatch, const std::vector<Keypoint>& ref_kp,
             const std::vector<Keypoint>& tar_kp, double thr);
    std::vector<float> gDist, sDist, gDist2, sDist2;
    std::vector<int> gIdx, sIdx, gIdx2, sIdx2;
    int status = 0;
};

// CSV of matched coordinates, one "x,y,z" per line (Src/cUtil.cc:938-954,1002-1016)
SIFT_LIBRARY_API void write_sift_kp(std::vector<Cvec>& kp, const char* file_name);
SIFT_LIBRARY_API void read_sift_kp(const char* file_name, std::vector<Cvec>& kp);

}  // namespace CPUSIFT

#endif
