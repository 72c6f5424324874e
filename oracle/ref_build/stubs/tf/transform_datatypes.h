// oracle/ref_build/stubs: stand-in for <tf/transform_datatypes.h> (TEST INFRASTRUCTURE; see ssf_ref_stubs.h)
#include "ssf_ref_stubs.h"
