// oracle/ref_build/stubs: stand-in for <std_msgs/Float64MultiArray.h> (TEST INFRASTRUCTURE; see ssf_ref_stubs.h)
#include "ssf_ref_stubs.h"
