// oracle/ref_build/stubs: stand-in for <tf/LinearMath/Quaternion.h> (TEST INFRASTRUCTURE; see ssf_ref_stubs.h)
#include "ssf_ref_stubs.h"
