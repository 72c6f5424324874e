"""B-wire: byte layouts of the two ROS1 messages the front end publishes, without ROS.

The C++ back end pairs message k on ``frame_odom1`` with the k-th plane cloud purely by arrival order
(``src/lidarOdometry.cpp:145-173``), so the contract is: exactly one odom per cloud, same order, these payloads:

* ``velodyne_points`` -- ``sensor_msgs/PointCloud2``: ``height=1, width=N``, fields x/y/z FLOAT32 at offsets 0/4/8,
  ``point_step=12``, ``row_step=12*N``, little endian, ``is_dense=False``, frame ``livox_frame``, data = N*12 bytes
  (``scripts/PointCloudOdometry.py:67-87``; the ASF drivers also *declare* an ``intensity`` field at offset 12 while
  keeping ``point_step=12`` -- ``ASF/main_sju_occ_ros.py:243-250`` -- reproduced by ``declare_intensity=True``).
* ``frame_odom1`` -- ``std_msgs/Float64MultiArray``: ``data = [tx,ty,tz,qx,qy,qz,qw]``
  (``scripts/PointCloudOdometry.py:97-103``; consumed at ``src/lidarOdometry.cpp:148-154`` as t then Quaterniond(x,y,z,w)).

``fill_ros_messages`` populates real rospy message objects when ROS is installed (not testable in this image);
``integrate_odometry`` replays the back end's pose integration (``src/lidarOdometry.cpp:80-81``) to emit TUM lines.
"""
import struct

import numpy as np

FLOAT32 = 7  # sensor_msgs/PointField.FLOAT32


def pointcloud2_fields(declare_intensity=False):
    f = [("x", 0, FLOAT32, 1), ("y", 4, FLOAT32, 1), ("z", 8, FLOAT32, 1)]
    if declare_intensity:
        f.append(("intensity", 12, FLOAT32, 1))
    return f


def pointcloud2_dict(points, declare_intensity=False, frame_id="livox_frame"):
    """points [N,3] -> dict with exactly the attributes the drivers set on sensor_msgs/PointCloud2."""
    p = np.ascontiguousarray(points, np.float32)
    assert p.ndim == 2 and p.shape[1] == 3
    return dict(frame_id=frame_id, height=1, width=p.shape[0], fields=pointcloud2_fields(declare_intensity),
                is_bigendian=False, point_step=12, row_step=12 * p.shape[0], is_dense=False, data=p.tobytes())


def odom_payload(odom7):
    """[tx,ty,tz,qx,qy,qz,qw] -> the 56 data bytes of the Float64MultiArray (little-endian float64)."""
    o = np.asarray(odom7, np.float64).reshape(7)
    return struct.pack("<7d", *o)


def serialize_float64_multiarray(odom7):
    """Full ROS1 wire serialisation of std_msgs/Float64MultiArray with an empty layout:
    uint32 dim count (0), uint32 data_offset (0), uint32 data length, then the float64s."""
    o = np.asarray(odom7, np.float64).reshape(-1)
    return struct.pack("<III", 0, 0, len(o)) + o.astype("<f8").tobytes()


def fill_ros_messages(points, odom7, stamp=None, declare_intensity=False):
    """Builds (PointCloud2, Float64MultiArray) rospy messages; requires a ROS1 Python environment."""
    from sensor_msgs.msg import PointCloud2, PointField  # noqa: E402  (only when ROS is present)
    from std_msgs.msg import Float64MultiArray
    d = pointcloud2_dict(points, declare_intensity)
    msg = PointCloud2()
    if stamp is not None:
        msg.header.stamp = stamp
    msg.header.frame_id = d["frame_id"]
    msg.height, msg.width = d["height"], d["width"]
    msg.fields = [PointField(n, off, dt, c) for n, off, dt, c in d["fields"]]
    msg.is_bigendian, msg.point_step, msg.row_step, msg.is_dense = False, 12, d["row_step"], False
    msg.data = d["data"]
    return msg, Float64MultiArray(data=list(np.asarray(odom7, np.float64)))


def _quat_mul(a, b):  # (x,y,z,w)
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def _quat_rot(q, v):
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    return R @ v


def integrate_odometry(odoms, dt=0.1):
    """Host replica of the back end's integration (src/lidarOdometry.cpp:80-81):
    q_0_curr = q_0_last * q_last_curr;  t_0_curr = t_0_last + q_0_last * t_last_curr.
    odoms [K,7] -> list of TUM lines 'time tx ty tz qx qy qz qw' (K+1 poses, first = identity)."""
    q0, t0 = np.array([0.0, 0, 0, 1]), np.zeros(3)
    lines = ["%.6f %.9f %.9f %.9f %.9f %.9f %.9f %.9f" % ((0.0,) + tuple(t0) + tuple(q0))]
    for k, o in enumerate(np.asarray(odoms, np.float64)):
        t_lc, q_lc = o[:3], o[3:]
        t0 = t0 + _quat_rot(q0, t_lc)
        q0 = _quat_mul(q0, q_lc)
        q0 = q0 / np.linalg.norm(q0)
        lines.append("%.6f %.9f %.9f %.9f %.9f %.9f %.9f %.9f" % (((k + 1) * dt,) + tuple(t0) + tuple(q0)))
    return lines
