"""B-wire: byte layouts of the two ROS1 messages the front end publishes, without ROS.

The C++ back end pairs message k on ``frame_odom1`` with the k-th plane cloud purely by arrival order
(``src/lidarOdometry.cpp:145-173``), so the contract is: exactly one odom per cloud, same order, these payloads:

* ``velodyne_points`` -- ``sensor_msgs/PointCloud2``: ``height=1, width=N``, fields x/y/z FLOAT32 at offsets 0/4/8,
  ``point_step=12``, ``row_step=12*N``, little endian, ``is_dense=False``, frame ``livox_frame``, data = N*12 bytes
  (``scripts/PointCloudOdometry.py:67-87``; the ASF drivers also *declare* an ``intensity`` field at offset 12 while
  keeping ``point_step=12`` -- ``ASF/main_sju_occ_ros.py:243-250`` -- reproduced by ``declare_intensity=True``).
* ``frame_odom1`` -- ``std_msgs/Float64MultiArray``: ``data = [tx,ty,tz,qx,qy,qz,qw]``
  (``scripts/PointCloudOdometry.py:97-103``; consumed at ``src/lidarOdometry.cpp:148-154`` as t then Quaterniond(x,y,z,w)).

``serialize_pointcloud2`` / ``serialize_float64_multiarray`` are the full ROS1 (genpy, little-endian) wire serialisations of
the two messages -- the bytes a subscriber's ``deserialize`` sees on the TCPROS connection after the 4-byte length prefix --
written from the message definitions (``sensor_msgs/PointCloud2.msg``, ``std_msgs/Header.msg``, ``sensor_msgs/PointField.msg``,
``std_msgs/Float64MultiArray.msg``, ``std_msgs/MultiArrayLayout.msg``); ``deserialize_*`` are their inverses.  ROS itself is not
in this image, so the byte goldens in tests/test_host.py are derived by hand from those definitions.
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


def _ros_string(text):
    b = text.encode("utf-8")
    return struct.pack("<I", len(b)) + b


def serialize_pointcloud2(points, seq=0, stamp=(0, 0), declare_intensity=False, frame_id="livox_frame"):
    """Full ROS1 wire bytes of the ``velodyne_points`` message for points [N,3] (fields, steps and flags as the reference
    drivers set them).  Layout, all little endian:
      Header   : uint32 seq | uint32 stamp.secs | uint32 stamp.nsecs | uint32 len + frame_id
      uint32 height | uint32 width
      fields[] : uint32 count, then per field: uint32 len + name | uint32 offset | uint8 datatype | uint32 count
      uint8 is_bigendian | uint32 point_step | uint32 row_step | uint32 len + data | uint8 is_dense"""
    d = pointcloud2_dict(points, declare_intensity, frame_id)
    return serialize_pointcloud2_dict(d, seq, stamp)


def serialize_pointcloud2_dict(d, seq=0, stamp=(0, 0)):
    out = [struct.pack("<III", seq, int(stamp[0]), int(stamp[1])), _ros_string(d["frame_id"]),
           struct.pack("<II", d["height"], d["width"]), struct.pack("<I", len(d["fields"]))]
    for name, offset, datatype, count in d["fields"]:
        out += [_ros_string(name), struct.pack("<IBI", offset, datatype, count)]
    out += [struct.pack("<BII", 1 if d["is_bigendian"] else 0, d["point_step"], d["row_step"]),
            struct.pack("<I", len(d["data"])), bytes(d["data"]), struct.pack("<B", 1 if d["is_dense"] else 0)]
    return b"".join(out)


def deserialize_pointcloud2(buf):
    """Inverse of serialize_pointcloud2 -> dict(seq, stamp, frame_id, height, width, fields, is_bigendian, point_step,
    row_step, data, is_dense); raises ValueError on trailing or missing bytes."""
    pos = 0

    def take(fmt):
        nonlocal pos
        v = struct.unpack_from(fmt, buf, pos)
        pos += struct.calcsize(fmt)
        return v

    def string():
        nonlocal pos
        (n,) = take("<I")
        v = bytes(buf[pos:pos + n])
        if len(v) != n:
            raise ValueError("truncated message")
        pos += n
        return v

    try:
        seq, secs, nsecs = take("<III")
        frame_id = string().decode("utf-8")
        height, width = take("<II")
        (nf,) = take("<I")
        fields = []
        for _ in range(nf):
            name = string().decode("utf-8")
            offset, datatype, count = take("<IBI")
            fields.append((name, offset, datatype, count))
        big, point_step, row_step = take("<BII")
        data = string()
        (dense,) = take("<B")
    except struct.error as e:
        raise ValueError("truncated message") from e
    if pos != len(buf):
        raise ValueError("%d trailing bytes" % (len(buf) - pos))
    return dict(seq=seq, stamp=(secs, nsecs), frame_id=frame_id, height=height, width=width, fields=fields,
                is_bigendian=bool(big), point_step=point_step, row_step=row_step, data=data, is_dense=bool(dense))


def deserialize_float64_multiarray(buf):
    """Inverse of serialize_float64_multiarray (any layout) -> (dims [(label, size, stride)], data_offset, float64 array)."""
    pos = 0
    (nd,) = struct.unpack_from("<I", buf, pos)
    pos += 4
    dims = []
    for _ in range(nd):
        (n,) = struct.unpack_from("<I", buf, pos)
        label = bytes(buf[pos + 4:pos + 4 + n]).decode("utf-8")
        pos += 4 + n
        size, stride = struct.unpack_from("<II", buf, pos)
        pos += 8
        dims.append((label, size, stride))
    data_offset, n = struct.unpack_from("<II", buf, pos)
    pos += 8
    if pos + 8 * n != len(buf):
        raise ValueError("length mismatch")
    return dims, data_offset, np.frombuffer(buf, "<f8", n, pos).copy()


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
