// oracle/ref_build/stubs: stand-in for <ros/ros.h> (TEST INFRASTRUCTURE; see ssf_ref_stubs.h)
#include "ssf_ref_stubs.h"
