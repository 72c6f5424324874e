// oracle/ref_build/stubs: stand-in for <sensor_msgs/PointCloud2.h> (TEST INFRASTRUCTURE; see ssf_ref_stubs.h)
#include "ssf_ref_stubs.h"
