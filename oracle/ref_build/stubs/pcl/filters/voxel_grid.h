// oracle/ref_build/stubs: stand-in for <pcl/filters/voxel_grid.h> (TEST INFRASTRUCTURE; see ssf_ref_stubs.h)
#include "ssf_ref_stubs.h"
