// oracle/ref_build/frame_feature_harness.cpp -- TEST INFRASTRUCTURE.
// C entry point around the reference's OWN compiled plane-feature node (src/frameFeature.cpp, built unmodified from
// /root/reference against oracle/ref_build/stubs): sets the node's globals as its main() does for a 16- or 64-line LiDAR
// (src/frameFeature.cpp:141-152), runs main() (which only registers the callback with the stub NodeHandle), hands the
// callback one cloud and returns what it published on /plane_frame_cloud1.
#include "ssf_ref_stubs.h"

extern int N_SCAN_ROW;                    // include/header.h:36 (a global, 16 by default)
extern float planeMin;                    // src/frameFeature.cpp:29-32
extern int planeSpan, rowIndexStart, rowIndexEnd;
int ssf_ref_frame_feature_main(int argc, char** argv);   // the node's main(), renamed on the compiler command line

extern "C" int ssf_ref_plane_features(const float* pts, int n, int n_rows, float* out_xyzi, int* out_count) {
    if (n_rows != 16 && n_rows != 64) return 1;
    N_SCAN_ROW = n_rows;
    planeMin = 0.5f; planeSpan = 2; rowIndexStart = 0; rowIndexEnd = 0;   // the file-scope initial values
    char arg0[] = "frame_feature";
    char* argv[] = {arg0, nullptr};
    ssf_ref_frame_feature_main(1, argv);  // sets planeMin / planeSpan / rowIndex* for N_SCAN_ROW and subscribes cloudHandler
    if (ssf_ref::handler() == nullptr) return 2;
    auto msg = std::make_shared<sensor_msgs::PointCloud2>();
    msg->xyzi.resize((size_t)n * 4);
    for (int i = 0; i < n; ++i) {
        msg->xyzi[4 * i] = pts[3 * i];
        msg->xyzi[4 * i + 1] = pts[3 * i + 1];
        msg->xyzi[4 * i + 2] = pts[3 * i + 2];
        msg->xyzi[4 * i + 3] = 0.f;       // the front end's cloud carries no intensity; the node overwrites it anyway (:77)
    }
    ssf_ref::published().clear();
    ssf_ref::handler()(msg);
    const sensor_msgs::PointCloud2& plane = ssf_ref::published()["/plane_frame_cloud1"];
    const int m = (int)(plane.xyzi.size() / 4);
    for (size_t i = 0; i < plane.xyzi.size(); ++i) out_xyzi[i] = plane.xyzi[i];
    *out_count = m;
    return 0;
}
