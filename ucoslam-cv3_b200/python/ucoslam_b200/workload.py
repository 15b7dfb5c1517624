"""Synthetic tracking workload of the BASELINE metric ("frames/sec (ORB+match+local-BA)"): n INDEPENDENT tracking problems cut from
one seeded flight over a textured plane (exact ground-truth poses; frames rendered by the exact homographies), so that a step of
the bench is what UcoSLAM does per frame and per keyframe on its hot path (src/utils/system.cpp:6460-6960, mapmanager.cpp):

  per frame    : ORB extraction -> search by projection from the previous frame -> solvePnp -> Map::matchFrameToMapPoints over the
                 local map -> solvePnp
  per keyframe : fbow transform (computeBow), FrameMatcher against the neighbouring keyframes (new-map-point creation), local BA

Problem f: the current image is frame f+2 of the flight; its "previous frame" is frame f+1 (keypoints + descriptors + one map point
per keypoint: its back-projection onto the plane, scale range as MapPoint::updateNormals sets it); the local map additionally holds
the points seen from frame f (not referenced by the previous frame: only the local-map search can find them); the pose prior is
the previous frame's pose.  The state is built with whatever `extract` callable is handed in (the CUDA extractor in the GPU arm,
the CPU restatement in the reference arm: they are bit-identical, so both arms track the same problems).  No oracle code here.
"""
import numpy as np

N_LEVELS, SCALE = 8, 1.2


class Camera:
    def __init__(self, w=640, h=480, f=525.0):
        self.w, self.h, self.f = w, h, float(f)
        self.cx, self.cy = w / 2 - 0.5, h / 2 - 0.5
        self.K = np.array([[self.f, 0, self.cx], [0, self.f, self.cy], [0, 0, 1.0]])
        self.px_per_m = 420.0 * (w / 640.0)      # texture pixels per metre on the plane z = 0


def _rodrigues(w):
    th = np.linalg.norm(w)
    Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + Kx
    return np.eye(3) + np.sin(th) / th * Kx + (1 - np.cos(th)) / th ** 2 * Kx @ Kx


def texture(seed=1234, size=2048):
    rng = np.random.default_rng(seed)
    acc, amp = np.zeros((size, size)), 1.0
    for blk in (64, 32, 16, 8, 4):
        n = size // blk
        acc += amp * np.kron(rng.random((n, n)), np.ones((blk, blk)))
        amp *= 0.5
    acc -= acc.min()
    return (acc / acc.max() * 255).astype(np.uint8)


def gt_pose(i, phase=0.0):
    """frame <- world, camera ~2 m above the plane z = 0 looking down the +z axis, on a smooth arc"""
    t = 0.04 * i + phase
    C = np.array([2.4 + 0.5 * np.sin(t), 2.4 + 0.35 * np.sin(1.3 * t), -2.0 + 0.15 * np.sin(0.7 * t)])
    R = _rodrigues(np.array([0.10 * np.sin(0.9 * t), 0.08 * np.sin(1.1 * t + 0.3), 0.12 * np.sin(0.5 * t)]))
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = -R @ C
    return T


def render(cam, tex, T):
    import cv2
    Hm = cam.K @ np.column_stack([T[:3, 0], T[:3, 1], T[:3, 3]]) @ np.diag([1 / cam.px_per_m, 1 / cam.px_per_m, 1.0])
    return cv2.warpPerspective(tex, Hm, (cam.w, cam.h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)


def back_project(cam, kps, T):
    """map points = a frame's keypoints back-projected onto the plane; scale-invariance range as MapPoint::updateNormals does
    (max = dist * scale^octave, min = max / scale^(levels-1)); normal = unit vector towards the observing camera"""
    R, t = T[:3, :3], T[:3, 3]
    C = -R.T @ t
    rays = (np.linalg.inv(cam.K) @ np.column_stack([kps["x"], kps["y"], np.ones(len(kps))]).T).T @ R
    lam = -C[2] / rays[:, 2]
    X = C + lam[:, None] * rays
    dist = np.linalg.norm(X - C, axis=1)
    sf = np.array([np.float32(SCALE) ** o for o in range(N_LEVELS)], np.float32)
    dmax = dist * sf[kps["octave"]]
    n = (C - X) / dist[:, None]
    return X.astype(np.float32), n.astype(np.float32), (dmax / sf[-1]).astype(np.float32), dmax.astype(np.float32)


def make_problems(n, extract, seed=1234, cam=None, phase=0.0):
    """-> (images (n, h, w) u8, list of n scene dicts with the keys of synth.synth_track_scene minus the current frame's keypoints,
    ground-truth poses (n, 4, 4) of the current frames)"""
    cam = cam or Camera()
    tex = texture(seed, 2048 if cam.w <= 640 else 4096)
    Ts = [gt_pose(i, phase) for i in range(n + 2)]
    imgs = np.stack([render(cam, tex, T) for T in Ts])
    feats = [extract(img) for img in imgs]
    sf = np.array([np.float32(SCALE) ** o for o in range(N_LEVELS)], np.float32)
    scenes = []
    for f in range(n):
        (k0, d0), (k1, d1) = feats[f], feats[f + 1]
        X1, n1, mn1, mx1 = back_project(cam, k1, Ts[f + 1])
        X0, n0, mn0, mx0 = back_project(cam, k0, Ts[f])
        m1, m0 = len(k1), len(k0)
        scenes.append(dict(
            prev_octave=k1["octave"].astype(np.int32), prev_desc=d1.copy(), prev_mp_row=np.arange(m1, dtype=np.int32),
            mp_id=(np.arange(m1 + m0, dtype=np.uint32) * 2 + 5), mp_pos=np.concatenate([X1, X0]), mp_normal=np.concatenate([n1, n0]),
            mp_min_dist=np.concatenate([mn1, mn0]), mp_max_dist=np.concatenate([mx1, mx0]), mp_desc=np.concatenate([d1, d0]),
            mp_stable=np.ones(m1 + m0, np.uint8), mp_local=np.ones(m1 + m0, np.uint8), scale_factors=sf,
            pose44=Ts[f + 1].astype(np.float32).reshape(16), fx=float(np.float32(cam.f)), fy=float(np.float32(cam.f)),
            cx=float(np.float32(cam.cx)), cy=float(np.float32(cam.cy)), bf=0.0, min_xy=np.array([0, 0], np.float32),
            max_xy=np.array([cam.w, cam.h], np.float32)))
    return imgs[2:].copy(), scenes, np.array(Ts[2:])


def with_current(sc, kps, desc):
    """a problem + the current frame's extraction = the scene dict the stage oracles / track_batch take"""
    return dict(sc, kp_xy=np.stack([kps["x"], kps["y"]], 1).astype(np.float32), kp_octave=kps["octave"].astype(np.int32), kp_desc=desc)


def synth_vocabulary_full(seed=7, k=10, depth=6, desc_size=32):
    """A seeded FULL k-ary vocabulary in fbow's stream format (3rdparty/fbow/fbow/fbow.cpp:160-190, fbow.h:125-194), breadth-first
    block numbering; k=10, depth=6 gives 111 111 blocks / 10^6 words — the size of the reference's shipped orb.fbow (110 259
    blocks).  Vectorised (the per-block generator of the tests takes minutes at this size)."""
    import struct
    rng = np.random.default_rng(seed)
    feature_off, desc_wp = 8, 32
    child_off = feature_off + k * desc_wp
    block_size = (child_off + 8 * k + 7) // 8 * 8
    nb = (k ** depth - 1) // (k - 1)
    n_internal = (k ** (depth - 1) - 1) // (k - 1)          # blocks whose children are blocks
    data = np.zeros((nb, block_size), np.uint8)
    data[:, 0:2] = np.frombuffer(np.uint16(k).tobytes(), np.uint8)
    parent = np.maximum(np.arange(nb, dtype=np.int64) - 1, 0) // k
    data[:, 4:8] = parent.astype("<u4").view(np.uint8).reshape(nb, 4)
    data[:, feature_off:feature_off + k * desc_wp] = rng.integers(0, 256, (nb, k * desc_wp), dtype=np.uint8)
    ids = np.zeros((nb, k), np.uint32)
    b = np.arange(nb, dtype=np.int64)[:, None]
    c = np.arange(k, dtype=np.int64)[None, :]
    ids[:n_internal] = (k * b[:n_internal] + 1 + c).astype(np.uint32)
    ids[n_internal:] = (0x80000000 | ((b[n_internal:] - n_internal) * k + c)).astype(np.uint32)
    w = (rng.random((nb, k)) * 3 + 0.01).astype(np.float32)
    w[:n_internal] = 0
    info = np.zeros((nb, k), dtype=[("id", "<u4"), ("w", "<f4")])
    info["id"], info["w"] = ids, w
    data[:, child_off:child_off + 8 * k] = info.view(np.uint8).reshape(nb, 8 * k)
    hdr = bytearray(128)
    struct.pack_into("<Q", hdr, 0, 55824124)
    hdr[8:11] = b"orb"
    struct.pack_into("<II", hdr, 8 + 52, 8, nb)
    struct.pack_into("<5Q", hdr, 8 + 64, desc_wp, block_size, feature_off, child_off, data.size)
    struct.pack_into("<iiI", hdr, 8 + 104, 0, desc_size, k)
    return np.concatenate([np.frombuffer(bytes(hdr), np.uint8), data.reshape(-1)])
