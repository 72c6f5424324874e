// oracle/ref_build/stubs: stand-in for <gtsam/geometry/Rot3.h> (TEST INFRASTRUCTURE; see ssf_ref_stubs.h)
#include "ssf_ref_stubs.h"
