/* Case shim: the reference includes "../Include/cutil.h" (cSIFT3D.cc:12) but ships cUtil.h.
 * Test infrastructure only; see oracle/README.md. */
#include "cUtil.h"
