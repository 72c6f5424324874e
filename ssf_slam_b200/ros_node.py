"""f-1: the reference's ROS front-end nodes as drop-in frame loops, runnable with or without ROS.

Each ``run_*`` function is the body of one reference driver's ``for`` loop, frame for frame:

* ``run_pointcloud_only``   -- scripts/PointCloudOdometry_onlyPC.py:36-66   (cloud only)
* ``run_gt_odometry``       -- scripts/PointCloudOdometry.py:59-105         (GT flow + GT mask ``s_fg_mask`` -> pose)
* ``run_noseg_odometry``    -- scripts/PointCloudOdometry_noSeg.py:62-127   (GT / stored flow + GMM mask -> pose)
* ``run_scene_flow_odometry`` -- ASF/main_sju_occ_ros.py:168-284            (network flow + mask -> pose; masker "gmm" is the
                                reference's, "residual" the north star's; Seg variants pass sem / inst labels)

Per frame they publish ONE ``velodyne_points`` cloud and then ONE ``frame_odom1`` payload, in that order: the C++ back end
pairs the k-th odometry with the k-th plane cloud purely by arrival order (src/lidarOdometry.cpp:145-173).  The mask / pose
work runs on the GPU through ``ssf_slam_b200.frontend``; only the message plumbing is host code.

``rospy`` is injected: pass the real module on a ROS machine, or ``Rosless()`` (below), an in-process stand-in with the same
call surface (``init_node``, ``Publisher(...).publish``, ``Rate(...).sleep``, ``Time.now``, ``has_param`` / ``get_param``,
``loginfo``, ``is_shutdown``) whose publishers record the ROS1 wire bytes of every message (``ssf_slam_b200.wire``), which is
what the tests inspect.  ``Rosless.msgs`` provides ``PointCloud2`` / ``PointField`` / ``Float64MultiArray`` stand-ins with the
attribute names of the genpy classes.
"""
import numpy as np

from . import wire


# ----------------------------------------------------------------------------------------------- message stand-ins
class _Time:
    def __init__(self, secs=0, nsecs=0):
        self.secs, self.nsecs = int(secs), int(nsecs)


class _Header:
    def __init__(self):
        self.seq, self.stamp, self.frame_id = 0, _Time(), ""


class PointField:
    INT8, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 = 1, 2, 3, 4, 5, 6, 7, 8

    def __init__(self, name="", offset=0, datatype=0, count=0):
        self.name, self.offset, self.datatype, self.count = name, offset, datatype, count


class PointCloud2:
    _has_header = True

    def __init__(self):
        self.header = _Header()
        self.height = self.width = self.point_step = self.row_step = 0
        self.fields, self.is_bigendian, self.is_dense, self.data = [], False, False, b""

    def serialize(self):
        d = dict(frame_id=self.header.frame_id, height=self.height, width=self.width,
                 fields=[(f.name, f.offset, f.datatype, f.count) for f in self.fields], is_bigendian=self.is_bigendian,
                 point_step=self.point_step, row_step=self.row_step, is_dense=self.is_dense, data=self.data)
        return wire.serialize_pointcloud2_dict(d, self.header.seq, (self.header.stamp.secs, self.header.stamp.nsecs))


class Float64MultiArray:
    _has_header = False

    def __init__(self, data=()):
        self.data = list(np.asarray(data, np.float64).reshape(-1))

    def serialize(self):
        return wire.serialize_float64_multiarray(self.data)


class _Msgs:
    PointCloud2, PointField, Float64MultiArray = PointCloud2, PointField, Float64MultiArray


class _Publisher:
    def __init__(self, owner, topic, data_class, queue_size):
        self.owner, self.topic, self.data_class, self.queue_size, self.seq = owner, topic, data_class, queue_size, 0

    def publish(self, msg):
        self.seq += 1                                  # rospy stamps header.seq on publish (rospy/topics.py)
        if getattr(msg, "_has_header", False):
            msg.header.seq = self.seq
        self.owner.log.append((self.topic, msg.serialize()))


class Rosless:
    """In-process stand-in for the ``rospy`` module surface the reference drivers use; ``log`` is the ordered list of
    (topic, ROS1 wire bytes) of everything published."""
    msgs = _Msgs

    def __init__(self, params=None, clock_step_ns=100_000_000):
        self.log, self.params, self.node, self._now, self._step, self.sleeps = [], dict(params or {}), None, 0, clock_step_ns, 0
        owner = self

        class Time(_Time):
            @staticmethod
            def now():
                return _Time(owner._now // 1_000_000_000, owner._now % 1_000_000_000)

        class Rate:
            def __init__(self, hz):
                self.hz = hz

            def sleep(self):            # simulated time: one tick per sleep, nothing blocks
                owner._now += owner._step
                owner.sleeps += 1

        self.Time, self.Rate = Time, Rate

    def init_node(self, name, anonymous=False):
        self.node = name

    def Publisher(self, topic, data_class, queue_size=10):
        return _Publisher(self, topic, data_class, queue_size)

    def has_param(self, key):
        return key in self.params

    def get_param(self, key, default=None):
        return self.params.get(key, default)

    def loginfo(self, *a):
        pass

    def is_shutdown(self):
        return False

    def topic(self, name):
        return [b for t, b in self.log if t == name]


# ----------------------------------------------------------------------------------------------- the frame loops
def _msg_types(rospy):
    if isinstance(rospy, Rosless):
        return rospy.msgs
    import sensor_msgs.msg as sm           # real ROS
    import std_msgs.msg as st

    class M:
        PointCloud2, PointField, Float64MultiArray = sm.PointCloud2, sm.PointField, st.Float64MultiArray
    return M


def _cloud_msg(rospy, M, points, declare_intensity):
    """The PointCloud2 the drivers fill field by field (scripts/PointCloudOdometry.py:67-87; the ASF drivers also declare an
    `intensity` field at offset 12 while keeping point_step = 12, ASF/main_sju_occ_ros.py:243-250)."""
    points = np.asarray(points)
    msg = M.PointCloud2()
    msg.header.stamp = rospy.Time.now()
    msg.header.frame_id = "livox_frame"
    if points.ndim == 3:
        msg.height, msg.width = points.shape[1], points.shape[0]
    else:
        msg.height, msg.width = 1, len(points)
    msg.fields = [M.PointField("x", 0, M.PointField.FLOAT32, 1), M.PointField("y", 4, M.PointField.FLOAT32, 1),
                  M.PointField("z", 8, M.PointField.FLOAT32, 1)]
    if declare_intensity:
        msg.fields.append(M.PointField("intensity", 12, M.PointField.FLOAT32, 1))
    msg.is_bigendian = False
    msg.point_step = 12
    msg.row_step = msg.point_step * points.shape[0]
    msg.is_dense = False
    msg.data = np.asarray(points, np.float32).tobytes()
    return msg


class _Node:
    def __init__(self, rospy, name, with_odom=True, queue_size=100):
        self.rospy, self.M = rospy, _msg_types(rospy)
        rospy.init_node(name, anonymous=True)
        self.cloud_pub = rospy.Publisher("velodyne_points", self.M.PointCloud2, queue_size=queue_size)
        self.odom_pub = rospy.Publisher("frame_odom1", self.M.Float64MultiArray, queue_size=100) if with_odom else None
        self.rate = rospy.Rate(10)  # 10 Hz

    def publish(self, points, odom7, declare_intensity=False):
        self.cloud_pub.publish(_cloud_msg(self.rospy, self.M, points, declare_intensity))
        if odom7 is not None:
            self.odom_pub.publish(self.M.Float64MultiArray(data=np.asarray(odom7, np.float64)))


def _frames(source):
    """source: iterable of npz paths (np.load'ed like the drivers do) or of dict-like frames."""
    for item in source:
        yield np.load(item) if isinstance(item, (str, bytes)) else item


def run_pointcloud_only(rospy, source):
    """scripts/PointCloudOdometry_onlyPC.py:36-66: publish `pos1` of every frame, nothing else."""
    node = _Node(rospy, "velodyne_points_node", with_odom=False, queue_size=10)
    n = 0
    for ac in _frames(source):
        if rospy.is_shutdown():
            break
        node.publish(ac["pos1"], None)
        node.rate.sleep()
        n += 1
    return n


def run_gt_odometry(rospy, source, device="cuda:0"):
    """scripts/PointCloudOdometry.py:59-105: cloud, then the pose of the points whose `s_fg_mask` is 0 under the GT flow."""
    from . import frontend
    node = _Node(rospy, "velodyne_points_odometry_node")
    odoms = []
    for ac in _frames(source):
        if rospy.is_shutdown():
            break
        out = frontend.odometry(ac["pos1"], ac["gt"], mask=ac["s_fg_mask"], device=device)
        node.publish(ac["pos1"], out["odom"])
        odoms.append(out["odom"])
        node.rate.sleep()
    return np.asarray(odoms)


def run_noseg_odometry(rospy, source, flow_key="gt", masker="gmm", device="cuda:0"):
    """scripts/PointCloudOdometry_noSeg.py:62-127: cloud, then GMM (majority component = background) on [flow | xyz] and the
    pose of the background points.  masker="residual" swaps in the north star's masker."""
    from . import frontend
    node = _Node(rospy, "velodyne_points_odometry_node")
    odoms = []
    for ac in _frames(source):
        if rospy.is_shutdown():
            break
        out = frontend.odometry(ac["pos1"], ac[flow_key], masker=masker, device=device)
        node.publish(ac["pos1"], out["odom"])
        odoms.append(out["odom"])
        node.rate.sleep()
    return np.asarray(odoms)


def run_scene_flow_odometry(rospy, front_end, source, seg=False, n_inst=None):
    """ASF/main_sju_occ_ros.py:168-284 (test_one_epoch with use_publish_ros): per frame pair the network's flow, the dynamic
    mask and the static-point pose on the GPU (`front_end` is a SceneFlowFrontEnd; its masker decides noSeg-GMM / residual);
    publishes the cloud with the `intensity` field declared as that driver does, then the odometry.  `seg=True` feeds the
    frame's `sem` / `inst` labels (Seg drivers).  Frames are processed one at a time, in order, like the reference's
    batch-size-1 loader; `front_end.submit` double-buffers the next frame's upload behind the current frame's kernels."""
    node = _Node(rospy, "velodyne_points_odometry_node")
    odoms = []
    n_slots = len(front_end._streams)
    pending = []          # (points, handle) in submission order

    def flush_one():
        pts, h = pending.pop(0)
        out = h.result()
        odom = out["odom"][0].numpy().copy()
        node.publish(pts, odom, declare_intensity=True)
        odoms.append(odom)
        node.rate.sleep()

    for k, ac in enumerate(_frames(source)):
        if rospy.is_shutdown():
            break
        kw = {}
        if seg:
            inst = np.asarray(ac["inst"], np.int32)
            kw = dict(sem=np.asarray(ac["sem"], np.int32)[None], inst=inst[None], n_inst=int(n_inst if n_inst is not None else inst.max() + 1))
        if len(pending) == n_slots:
            flush_one()     # the slot about to be reused must have delivered (its pinned buffers are overwritten by submit)
        h = front_end.submit(np.asarray(ac["pos1"], np.float32)[None], np.asarray(ac["pos2"], np.float32)[None], slot=k % n_slots, **kw)
        pending.append((np.asarray(ac["pos1"]), h))
    while pending:
        flush_one()
    return np.asarray(odoms)
