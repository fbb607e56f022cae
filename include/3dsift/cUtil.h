// 3dsift/cUtil.h — the exported part of the reference's Include/cUtil.h (:66-68): the CSV consumers of a match
// (SURVEY.md section 8f-3).  Everything else in that header is internal to the reference's CPU pipeline (not exported).
#ifndef S3D_FACADE_CUTIL_H
#define S3D_FACADE_CUTIL_H

#include <vector>

#include "Util/common.h"
#include "cMatcher.h"  // declares write_sift_kp / read_sift_kp next to the matcher whose output they store
#include "cSIFT3D.h"

#endif
