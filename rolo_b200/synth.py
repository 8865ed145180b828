"""Synthetic LiDAR scenes and scans (SURVEY.md §8d "Synthetic inputs").

Analytic scene, no files: ground plane z=-1.8 m, a 60x40x8 m room centred at (5,3), 12 vertical
cylinders (r=0.4 m) and 6 axis-aligned boxes at seeded positions, ray-cast from a sensor pose.
Sensor models: Velodyne-64 style (64 rings, -24.8..+2.0 deg, 1024 columns) for config C1 and
OS1-128 style (128 rings, +-22.5 deg, 2048 columns) for config C2. Range noise N(0, sigma).
The sensor always sits inside the closed room, so every ray hits: clouds have exactly
rings*cols points (ring-major order, the order a driver delivers them in).

Pure numpy; used by tests and bench.py to build inputs. Nothing here touches the oracle.
"""
from __future__ import annotations

import numpy as np

ROOM_MIN = np.array([5.0 - 30.0, 3.0 - 20.0, -1.8])
ROOM_MAX = np.array([5.0 + 30.0, 3.0 + 20.0, -1.8 + 8.0])

SENSORS = {
    # name: (rings, cols, elev_min_deg, elev_max_deg)
    "vlp64": (64, 1024, -24.8, 2.0),
    "os1-128": (128, 2048, -22.5, 22.5),
    "os1-128x1024": (128, 1024, -22.5, 22.5),
    "tiny": (16, 256, -15.0, 15.0),
    "small": (32, 512, -20.0, 10.0),
}


def rot_rpy(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """R = Rz(yaw) @ Ry(pitch) @ Rx(roll), radians."""
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def pose(rpy_deg=(0.0, 0.0, 0.0), t=(0.0, 0.0, 0.0)) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = rot_rpy(*np.deg2rad(np.asarray(rpy_deg, dtype=np.float64)))
    T[:3, 3] = np.asarray(t, dtype=np.float64)
    return T


#: BASELINE.json config 1/2 offset: rpy=(+2,-3,+5) deg, t=(0.20,-0.10,0.05) m
T_GT_DEFAULT = pose((2.0, -3.0, 5.0), (0.20, -0.10, 0.05))


def make_scene(seed: int = 7) -> dict:
    rng = np.random.default_rng(seed)
    cyl = []
    while len(cyl) < 12:
        c = np.array([rng.uniform(ROOM_MIN[0] + 2, ROOM_MAX[0] - 2), rng.uniform(ROOM_MIN[1] + 2, ROOM_MAX[1] - 2)])
        if np.hypot(*c) > 4.0:
            cyl.append(c)
    boxes = []
    while len(boxes) < 6:
        c = np.array([rng.uniform(ROOM_MIN[0] + 4, ROOM_MAX[0] - 4), rng.uniform(ROOM_MIN[1] + 4, ROOM_MAX[1] - 4)])
        half = rng.uniform(0.5, 2.0, size=2)
        h = rng.uniform(0.8, 3.0)
        if np.hypot(*c) > 6.0:
            boxes.append((np.array([c[0] - half[0], c[1] - half[1], ROOM_MIN[2]]), np.array([c[0] + half[0], c[1] + half[1], ROOM_MIN[2] + h])))
    return {"cylinders": np.array(cyl), "cyl_r": 0.4, "boxes": boxes}


def _slab(o, d, bmin, bmax):
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        t0 = (bmin - o) * inv
        t1 = (bmax - o) * inv
    tn = np.minimum(t0, t1)
    tf = np.maximum(t0, t1)
    tn = np.where(np.isnan(tn), -np.inf, tn)
    tf = np.where(np.isnan(tf), np.inf, tf)
    return tn.max(axis=1), tf.min(axis=1)


def raycast(scene: dict, T_ws: np.ndarray, dirs_s: np.ndarray) -> np.ndarray:
    """Distance along each sensor-frame unit ray (N,3) to the first surface."""
    o = T_ws[:3, 3][None, :]
    d = dirs_s @ T_ws[:3, :3].T
    n = d.shape[0]
    # inside the room: exit distance of the slab test is the wall/ceiling/floor hit
    _, tf = _slab(o, d, ROOM_MIN[None, :], ROOM_MAX[None, :])
    t = tf.copy()
    # boxes: entry distance
    for bmin, bmax in scene["boxes"]:
        tn, tf2 = _slab(o, d, bmin[None, :], bmax[None, :])
        hit = (tn <= tf2) & (tn > 1e-6)
        t = np.where(hit & (tn < t), tn, t)
    # vertical cylinders (infinite in z; the room bounds them)
    r = scene["cyl_r"]
    a = d[:, 0] ** 2 + d[:, 1] ** 2
    for c in scene["cylinders"]:
        oc = o[0, :2] - c
        b = 2.0 * (d[:, 0] * oc[0] + d[:, 1] * oc[1])
        cc = oc @ oc - r * r
        disc = b * b - 4.0 * a * cc
        with np.errstate(divide="ignore", invalid="ignore"):
            tc = (-b - np.sqrt(np.maximum(disc, 0.0))) / (2.0 * a)
        hit = (disc > 0) & (a > 1e-12) & (tc > 1e-6)
        t = np.where(hit & (tc < t), tc, t)
    assert t.shape == (n,)
    return t


def ray_dirs(sensor: str) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    rings, cols, e0, e1 = SENSORS[sensor]
    elev = np.deg2rad(np.linspace(e0, e1, rings))
    # azimuth of column c sits at the centre of imageProjection's column c (imageProjection.cpp:438-444)
    az = -np.pi + 2.0 * np.pi * np.arange(cols) / cols
    ee, aa = np.meshgrid(elev, az, indexing="ij")
    d = np.stack([np.cos(ee) * np.cos(aa), np.cos(ee) * np.sin(aa), np.sin(ee)], axis=-1).reshape(-1, 3)
    ring = np.repeat(np.arange(rings), cols)
    col = np.tile(np.arange(cols), rings)
    return d, ring, col


def scan(scene: dict, T_ws: np.ndarray, sensor: str = "os1-128", sigma: float = 0.02, seed: int = 1234,
         return_meta: bool = False):
    """One LiDAR sweep from sensor pose T_ws; returns float32 (N,4) xyzw (w=1) in the sensor frame."""
    d, ring, col = ray_dirs(sensor)
    t = raycast(scene, T_ws, d)
    rng = np.random.default_rng(seed)
    t = t + rng.normal(0.0, sigma, size=t.shape)
    pts = np.empty((d.shape[0], 4), dtype=np.float32)
    pts[:, :3] = (d * t[:, None]).astype(np.float32)
    pts[:, 3] = 1.0
    if return_meta:
        return pts, ring.astype(np.int32), col.astype(np.int32)
    return pts


def scan_pair(sensor: str = "os1-128", T_gt: np.ndarray | None = None, seed_target: int = 1234, seed_source: int = 5678,
              scene_seed: int = 7, sigma: float = 0.02, T_base: np.ndarray | None = None):
    """(target, source, T_gt): target scanned at T_base (default identity), source at T_base @ T_gt.
    Registration of source onto target recovers T_gt."""
    scene = make_scene(scene_seed)
    if T_gt is None:
        T_gt = T_GT_DEFAULT
    if T_base is None:
        T_base = np.eye(4)
    target = scan(scene, T_base, sensor, sigma, seed_target)
    source = scan(scene, T_base @ T_gt, sensor, sigma, seed_source)
    return target, source, T_gt


def random_offset(rng: np.random.Generator, max_deg: float = 5.0, max_t: float = 0.2) -> np.ndarray:
    """Offset drawn U(+-max_deg, +-max_t) (config C3)."""
    return pose(rng.uniform(-max_deg, max_deg, 3), rng.uniform(-max_t, max_t, 3))


def pose_error(T_a: np.ndarray, T_b: np.ndarray) -> tuple[float, float]:
    """(rotation angle [rad], translation distance [m]) between two poses."""
    D = np.linalg.inv(T_a) @ T_b
    # atan2(|sin|, cos) from the antisymmetric part: first-order accurate for small angles. arccos of the trace alone turns the
    # 6e-8 rounding of a float32 rotation matrix into an apparent 3e-4 rad.
    c = (np.trace(D[:3, :3]) - 1.0) / 2.0
    s = 0.5 * np.linalg.norm([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]])
    return float(np.arctan2(s, c)), float(np.linalg.norm(T_a[:3, 3] - T_b[:3, 3]))


# ---- torch variants (used by bench.py to build many full-size scans quickly on the GPU) --------------
def scan_torch(scene: dict, T_ws: np.ndarray, sensor: str = "os1-128", sigma: float = 0.02, seed: int = 1234, device="cuda"):
    """Same geometry as scan(), evaluated with torch in fp64 on `device`; noise from a torch generator.
    Returns a float32 (N,4) tensor on `device` (w = 1). Not bit-identical to scan() (different RNG)."""
    import torch
    d_np, _, _ = ray_dirs(sensor)
    dt = torch.float64
    d_s = torch.as_tensor(d_np, dtype=dt, device=device)
    R = torch.as_tensor(T_ws[:3, :3], dtype=dt, device=device)
    o = torch.as_tensor(T_ws[:3, 3], dtype=dt, device=device)
    d = d_s @ R.T
    inf = torch.tensor(float("inf"), dtype=dt, device=device)

    def slab(bmin, bmax):
        bmin = torch.as_tensor(bmin, dtype=dt, device=device); bmax = torch.as_tensor(bmax, dtype=dt, device=device)
        inv = 1.0 / d
        t0 = (bmin - o) * inv; t1 = (bmax - o) * inv
        tn = torch.minimum(t0, t1); tf = torch.maximum(t0, t1)
        tn = torch.where(torch.isnan(tn), -inf, tn); tf = torch.where(torch.isnan(tf), inf, tf)
        return tn.max(dim=1).values, tf.min(dim=1).values

    _, t = slab(ROOM_MIN, ROOM_MAX)
    for bmin, bmax in scene["boxes"]:
        tn, tf2 = slab(bmin, bmax)
        hit = (tn <= tf2) & (tn > 1e-6) & (tn < t)
        t = torch.where(hit, tn, t)
    r = scene["cyl_r"]
    a = d[:, 0] ** 2 + d[:, 1] ** 2
    for c in scene["cylinders"]:
        oc0 = float(T_ws[0, 3] - c[0]); oc1 = float(T_ws[1, 3] - c[1])
        b = 2.0 * (d[:, 0] * oc0 + d[:, 1] * oc1)
        cc = oc0 * oc0 + oc1 * oc1 - r * r
        disc = b * b - 4.0 * a * cc
        tc = (-b - torch.sqrt(torch.clamp(disc, min=0.0))) / (2.0 * a)
        hit = (disc > 0) & (a > 1e-12) & (tc > 1e-6) & (tc < t)
        t = torch.where(hit, tc, t)
    g = torch.Generator(device=device); g.manual_seed(int(seed))
    t = t + sigma * torch.randn(t.shape, dtype=dt, device=device, generator=g)
    pts = torch.ones((d_s.shape[0], 4), dtype=torch.float32, device=device)
    pts[:, :3] = (d_s * t[:, None]).to(torch.float32)
    return pts


def raw_frame(scene: dict, T_ws: np.ndarray, sensor: str = "os1-128", sigma: float = 0.02, seed: int = 1234, extra: int = 0) -> np.ndarray:
    """A raw driver frame in the reference's VelodynePointXYZIRT layout (imageProjection.cpp:7-19), points in firing
    order (all rings of azimuth step 0, then step 1, ...). `extra` appends jittered duplicates of random points so that
    several points compete for one range-image cell (exercises "first point wins", imageProjection.cpp:451)."""
    from .frontend import RAW_POINT_DTYPE
    pts, ring, col = scan(scene, T_ws, sensor, sigma, seed, return_meta=True)
    rings, cols, _, _ = SENSORS[sensor]
    order = np.lexsort((ring, col))
    pts, ring, col = pts[order], ring[order], col[order]
    if extra:
        rng = np.random.default_rng(seed + 1)
        pick = rng.integers(0, pts.shape[0], size=extra)
        dup = pts[pick].copy()
        dup[:, :3] *= (1.0 + rng.normal(0.0, 2e-3, size=(extra, 1))).astype(np.float32)
        where = np.sort(rng.integers(0, pts.shape[0], size=extra))
        pts = np.insert(pts, where, dup, axis=0); ring = np.insert(ring, where, ring[pick]); col = np.insert(col, where, col[pick])
    raw = np.zeros(pts.shape[0], dtype=RAW_POINT_DTYPE)
    raw["x"], raw["y"], raw["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    raw["intensity"] = 1.0
    raw["ring"] = ring.astype(np.uint16)
    raw["time"] = (col.astype(np.float32) / cols) * 0.1
    return raw


# ---- scan-to-map cases (backMapping.cpp:681-1054) -----------------------------------------------------------------

def voxel_downsample(pts: np.ndarray, leaf: float) -> np.ndarray:
    """Centroid per occupied voxel (what pcl::VoxelGrid computes); only used to build synthetic inputs."""
    if pts.shape[0] == 0:
        return pts.reshape(0, 4).astype(np.float32)
    key = np.floor(pts[:, :3].astype(np.float64) / leaf).astype(np.int64)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    out = np.zeros((cnt.shape[0], 4), dtype=np.float64)
    np.add.at(out, inv, pts[:, :4].astype(np.float64))
    out /= cnt[:, None]
    return out.astype(np.float32)


def pose6(T: np.ndarray) -> np.ndarray:
    """4x4 -> (roll, pitch, yaw, x, y, z) float32, the reference's transformTobeMapped order (ZYX Euler angles)."""
    R = T[:3, :3]
    pitch = -np.arcsin(np.clip(R[2, 0], -1.0, 1.0))
    roll = np.arctan2(R[2, 1], R[2, 2]); yaw = np.arctan2(R[1, 0], R[0, 0])
    return np.array([roll, pitch, yaw, T[0, 3], T[1, 3], T[2, 3]], dtype=np.float32)


def pose6_matrix(tf: np.ndarray) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = rot_rpy(float(tf[0]), float(tf[1]), float(tf[2])); T[:3, 3] = tf[3:6]
    return T


def scan2map_case(features, sensor: str = "small", n_keyframes: int = 6, seed: int = 11, corner_leaf: float = 0.2, surf_leaf: float = 0.4,
                  guess_deg: float = 1.0, guess_t: float = 0.15) -> dict:
    """A local feature map assembled from `n_keyframes` keyframes (extractSurroundingKeyFrames' product: world-frame
    corner/surface clouds, voxel-downsampled) and the downsampled features of a new scan with an offset initial guess.
    `features(raw) -> (corner (n,4), surface (n,4))` is the front end to use (the GPU one or the oracle's)."""
    scene = make_scene()
    rng = np.random.default_rng(seed)
    poses = [pose((0.3 * k, -0.2 * k, 4.0 * k), (0.8 * k, 0.15 * k, 0.0)) for k in range(n_keyframes + 1)]
    mc, ms = [], []
    for k in range(n_keyframes):
        c, s = features(raw_frame(scene, poses[k], sensor, seed=seed + k))
        for dst, f in ((mc, c), (ms, s)):
            w = f.astype(np.float64).copy(); w[:, :3] = f[:, :3].astype(np.float64) @ poses[k][:3, :3].T + poses[k][:3, 3]
            dst.append(w.astype(np.float32))
    c, s = features(raw_frame(scene, poses[-1], sensor, seed=seed + 100))
    T_true = poses[-1]
    T_guess = T_true @ pose((guess_deg * rng.uniform(-1, 1), guess_deg * rng.uniform(-1, 1), guess_deg * rng.uniform(-1, 1)),
                            tuple(guess_t * rng.uniform(-1, 1, 3)))
    return {"map_corner": voxel_downsample(np.concatenate(mc), corner_leaf), "map_surf": voxel_downsample(np.concatenate(ms), surf_leaf),
            "scan_corner": voxel_downsample(c, corner_leaf), "scan_surf": voxel_downsample(s, surf_leaf),
            "pose_true": pose6(T_true), "pose_guess": pose6(T_guess), "T_true": T_true}
