"""Synthetic CARLA-shaped LiDAR frame pairs (host side, NumPy).

There is no dataset in the build or GPU environment, so bench / tests use procedurally generated
clouds shaped like the reference's ``rm_road/SF/xx/*.npz`` items: a 64-channel, 100 m LiDAR mounted at
z = 2.5 m (reference sensor: ``ASF/Scenario_Traj.py:308-311,437``), road returns removed, random
subsampling to exactly N points with ``replace = (available < N)`` as the reference loader does
(``ASF/utils/datasets/carla.py:274-285`` -- this creates duplicate points on purpose), and the npz
key set ``pos1,pos2,gt,ego_flow,s_fg_mask,t_fg_mask`` (``ASF/utils/datasets/carla.py:432-488``) plus
synthetic ``sem`` / ``inst`` labels for the Seg configuration.

Scene: a straight street, two building facades, poles, 8-20 box vehicles of which 30-60 % move.
Frame t+1 is ray-cast independently, so ``pos2`` has no point correspondence with ``pos1``.
Seeds follow SURVEY.md section 8(d).
"""
import numpy as np

SEM_BUILDING, SEM_POLE, SEM_VEHICLE = 0, 1, 2
MOVABLE_CLASSES = (SEM_VEHICLE,)


def _rot_z(yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


class Street:
    """World-frame scene + per-frame ego and vehicle poses for one trajectory."""

    def __init__(self, seed, n_frames):
        rng = np.random.default_rng(seed)
        self.n_frames = n_frames
        self.half_width = rng.uniform(9.0, 14.0)
        self.wall_h = rng.uniform(8.0, 20.0)
        n_pole = int(rng.integers(10, 24))
        self.poles = np.stack([rng.uniform(-60, 260, n_pole),
                               rng.choice([-1.0, 1.0], n_pole) * (self.half_width - rng.uniform(1.0, 2.5, n_pole)),
                               rng.uniform(4.0, 8.0, n_pole)], 1)  # x, y, height
        n_veh = int(rng.integers(8, 21))
        self.veh_xy0 = np.stack([rng.uniform(-40, 120, n_veh), rng.uniform(-self.half_width + 2.5, self.half_width - 2.5, n_veh)], 1)
        self.veh_yaw = rng.normal(0.0, 0.05, n_veh) + rng.choice([0.0, np.pi], n_veh)
        moving = rng.random(n_veh) < rng.uniform(0.3, 0.6)
        self.veh_speed = np.where(moving, rng.uniform(0.0, 2.0, n_veh), 0.0)  # m / frame along heading
        self.veh_size = np.array([4.5, 1.9, 1.6])
        # ego: 0.3-1.5 m/frame forward, yaw rate <= 2 deg/frame
        v = rng.uniform(0.3, 1.5) + 0.1 * rng.standard_normal(n_frames + 1)
        v = np.clip(v, 0.3, 1.5)
        dyaw = np.deg2rad(np.clip(0.5 * rng.standard_normal(n_frames + 1), -2.0, 2.0)) * 0.3
        yaw = np.cumsum(dyaw) - dyaw[0]
        xy = np.zeros((n_frames + 1, 2))
        for i in range(1, n_frames + 1):
            xy[i] = xy[i - 1] + v[i] * np.array([np.cos(yaw[i - 1]), np.sin(yaw[i - 1])])
        self.ego_xy, self.ego_yaw = xy, yaw

    def ego_pose(self, f):
        """(R, t): sensor -> world."""
        return _rot_z(self.ego_yaw[f]), np.array([self.ego_xy[f, 0], self.ego_xy[f, 1], 2.5])

    def veh_pose(self, f):
        heading = np.stack([np.cos(self.veh_yaw), np.sin(self.veh_yaw)], 1)
        return self.veh_xy0 + heading * (self.veh_speed * f)[:, None]


def _ray_dirs():
    elev = np.deg2rad(np.linspace(-25.0, 5.0, 64))
    azim = np.linspace(-np.pi, np.pi, 1024, endpoint=False)
    e, a = np.meshgrid(elev, azim, indexing="ij")
    return np.stack([np.cos(e) * np.cos(a), np.cos(e) * np.sin(a), np.sin(e)], -1).reshape(-1, 3)


_DIRS = None


def _slab(o, d, lo, hi):
    """Ray/AABB: o [R,3], d [R,3], lo/hi [3] -> t_hit [R] (inf if miss)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        t0 = (lo - o) * inv
        t1 = (hi - o) * inv
    tmin = np.nanmax(np.minimum(t0, t1), axis=1)
    tmax = np.nanmin(np.maximum(t0, t1), axis=1)
    hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0.05)
    return np.where(hit, tmin, np.inf)


def _scan(scene, f, rng):
    """Ray-cast frame f.  Returns sensor-frame points [M,3] f64, sem [M], inst [M], vehicle id [M] (-1 = none)."""
    global _DIRS
    if _DIRS is None:
        _DIRS = _ray_dirs()
    R, t = scene.ego_pose(f)
    d = _DIRS @ R.T
    o = np.broadcast_to(t, d.shape)
    nr = d.shape[0]
    best = np.full(nr, np.inf)
    sem = np.full(nr, -1, np.int32)
    inst = np.zeros(nr, np.int32)
    veh = np.full(nr, -1, np.int32)

    def take(th, s, i, v, rows=None):
        if rows is None:
            m = th < best
            best[m] = th[m]
            sem[m] = s
            inst[m] = i
            veh[m] = v
        else:   # th holds the rays `rows` only
            m = th < best[rows]
            r = rows[m]
            best[r] = th[m]
            sem[r] = s
            inst[r] = i
            veh[r] = v

    az = np.arctan2(d[:, 1], d[:, 0])

    def rays_towards(corners_xy):
        """Indices of the rays whose azimuth lies inside the angle the footprint subtends (plus a margin); None = all rays
        (sensor inside or next to the footprint).  Small objects are hit by a few percent of the 65536 rays, so only those
        are intersected."""
        rel = corners_xy - t[:2]
        if np.min(np.hypot(rel[:, 0], rel[:, 1])) < 0.5 or (rel[:, 0].min() < 0 < rel[:, 0].max() and rel[:, 1].min() < 0 < rel[:, 1].max()):
            return None
        ang = np.arctan2(rel[:, 1], rel[:, 0])
        c = np.arctan2(rel[:, 1].mean(), rel[:, 0].mean())
        dif = (ang - c + np.pi) % (2 * np.pi) - np.pi
        if dif.max() - dif.min() > np.pi / 2:
            return None
        da = (az - c + np.pi) % (2 * np.pi) - np.pi
        return np.flatnonzero((da >= dif.min() - 0.02) & (da <= dif.max() + 0.02))

    # ground (z = 0): blocks rays, later dropped ("rm_road")
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = np.where(d[:, 2] < 0, -t[2] / d[:, 2], np.inf)
    take(tg, -2, 0, -1)
    for sgn in (-1.0, 1.0):
        y = sgn * scene.half_width
        take(_slab(o, d, np.array([-100.0, min(y, y + sgn), 0.0]), np.array([400.0, max(y, y + sgn), scene.wall_h])),
             SEM_BUILDING, 0, -1)
    sq = np.array([[-1.0, -1.0], [-1.0, 1.0], [1.0, -1.0], [1.0, 1.0]])
    for (px, py, ph) in scene.poles:
        rows = rays_towards(np.array([px, py]) + 0.15 * sq)
        lo, hi = np.array([px - 0.15, py - 0.15, 0.0]), np.array([px + 0.15, py + 0.15, ph])
        if rows is None:
            take(_slab(o, d, lo, hi), SEM_POLE, 0, -1)
        elif len(rows):
            take(_slab(o[rows], d[rows], lo, hi), SEM_POLE, 0, -1, rows)
    vxy = scene.veh_pose(f)
    half = scene.veh_size / 2
    for k in range(len(vxy)):
        Rv = _rot_z(scene.veh_yaw[k])
        c = np.array([vxy[k, 0], vxy[k, 1], half[2]])
        rows = rays_towards(vxy[k] + (sq * half[:2]) @ Rv[:2, :2].T)
        if rows is None:
            take(_slab((o - c) @ Rv, d @ Rv, -half, half), SEM_VEHICLE, k + 1, k)
        elif len(rows):
            take(_slab((o[rows] - c) @ Rv, d[rows] @ Rv, -half, half), SEM_VEHICLE, k + 1, k, rows)
    ok = np.isfinite(best) & (best <= 100.0) & (sem >= 0)
    noise = 0.01 * rng.standard_normal(nr)
    pts = _DIRS[ok] * (best[ok] + noise[ok])[:, None]
    return pts, sem[ok], inst[ok], veh[ok]


def _subsample(rng, m, n):
    return rng.choice(m, n, replace=(m < n))


def frame_pair(scene, f, n_points, rng, scan1=None, scan2=None):
    """One npz-shaped item: dict(pos1,pos2,gt,ego_flow,s_fg_mask,t_fg_mask,sem,inst), fp32 / int."""
    p1, sem1, inst1, veh1 = scan1 if scan1 is not None else _scan(scene, f, rng)
    p2, _, _, veh2 = scan2 if scan2 is not None else _scan(scene, f + 1, rng)
    i1 = _subsample(rng, len(p1), n_points)
    i2 = _subsample(rng, len(p2), n_points)
    p1, sem1, inst1, veh1 = p1[i1], sem1[i1], inst1[i1], veh1[i1]
    p2, veh2 = p2[i2], veh2[i2]
    R0, t0 = scene.ego_pose(f)
    R1, t1 = scene.ego_pose(f + 1)
    world = p1 @ R0.T + t0
    ego_flow = (world - t1) @ R1 - p1  # static world seen from the next sensor pose
    moved = world.copy()
    heading = np.stack([np.cos(scene.veh_yaw), np.sin(scene.veh_yaw)], 1)
    on = veh1 >= 0
    moved[on, :2] += heading[veh1[on]] * scene.veh_speed[veh1[on]][:, None]
    gt = (moved - t1) @ R1 - p1
    fg1 = on & (scene.veh_speed[np.maximum(veh1, 0)] > 1e-3)
    fg2 = (veh2 >= 0) & (scene.veh_speed[np.maximum(veh2, 0)] > 1e-3)
    return dict(pos1=p1.astype(np.float32), pos2=p2.astype(np.float32), gt=gt.astype(np.float32),
                ego_flow=ego_flow.astype(np.float32), s_fg_mask=fg1.astype(np.uint8), t_fg_mask=fg2.astype(np.uint8),
                sem=sem1.astype(np.int32), inst=inst1.astype(np.int32))


def make_sequence(seed, n_frames, n_points):
    """List of n_frames frame-pair dicts along one trajectory (BASELINE.json configs 2-4).  Every frame is
    ray-cast once; pair f uses scan f (subsampled) as pos1 and scan f+1 (subsampled independently) as pos2."""
    scene = Street(seed, n_frames)
    rng = np.random.default_rng(seed + 7919)
    scans = {}

    def scan(f):
        if f not in scans:
            scans[f] = _scan(scene, f, rng)
        return scans[f]

    out = []
    for f in range(n_frames):
        out.append(frame_pair(scene, f, n_points, rng, scan(f), scan(f + 1)))
        scans.pop(f, None)
    return out


def make_pair(seed, n_points):
    """A single frame pair (BASELINE.json config 1: seed 0)."""
    return make_sequence(seed, 1, n_points)[0]


def dense_cloud(seed, n_points):
    """Uniform-density full-beam stress cloud for the radius sweep (config 5): [n,3] f32."""
    rng = np.random.default_rng(seed)
    r = 100.0 * np.sqrt(rng.random(n_points))
    a = rng.uniform(-np.pi, np.pi, n_points)
    z = rng.uniform(-2.5, 12.0, n_points)
    return np.stack([r * np.cos(a), r * np.sin(a), z], 1).astype(np.float32)
