"""A minimal tracking chain over the hot path, for the 'matched ATE' check of the BASELINE metric: a camera flies over a textured
plane (exact ground-truth poses, frames rendered by the exact homographies), a map is made from the first frame's keypoints, and
every following frame is tracked as the reference's tracker does on its hot path:
    ORB extract (a1-a8) -> Map::matchFrameToMapPoints against the map with the previous pose (a12) -> PnPSolver::solvePnp (a21).
The chain takes its three stages as callables, so the SAME driver runs the CUDA path (through the C ABI) and, in tests / scripts,
the CPU oracle; both trajectories and their absolute trajectory errors (RMSE of the camera centres against ground truth) are
compared.  No oracle code is imported here."""
import numpy as np

W, H, F = 640, 480, 525.0
CX, CY = W / 2 - 0.5, H / 2 - 0.5
K = np.array([[F, 0, CX], [0, F, CY], [0, 0, 1.0]])
N_LEVELS, SCALE = 8, 1.2
PX_PER_M = 420.0          # texture pixels per metre on the plane z = 0


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


def gt_pose(i):
    """frame <- world, camera ~2 m above the plane z = 0 looking down the +z axis, on a smooth arc"""
    t = 0.04 * i
    C = np.array([2.4 + 0.5 * np.sin(t), 2.4 + 0.35 * np.sin(1.3 * t), -2.0 + 0.15 * np.sin(0.7 * t)])
    R = _rodrigues(np.array([0.10 * np.sin(0.9 * t), 0.08 * np.sin(1.1 * t + 0.3), 0.12 * np.sin(0.5 * t)]))
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = -R @ C
    return T


def render(tex, T):
    import cv2
    Hm = K @ np.column_stack([T[:3, 0], T[:3, 1], T[:3, 3]]) @ np.diag([1 / PX_PER_M, 1 / PX_PER_M, 1.0])
    return cv2.warpPerspective(tex, Hm, (W, H), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)


def make_map(kps, desc, T):
    """map points = the first frame's keypoints back-projected onto the plane; scale-invariance range as the reference's
    MapPoint::updateNormals does (max = dist * scale^octave, min = max / scale^(levels-1))"""
    R, t = T[:3, :3], T[:3, 3]
    C = -R.T @ t
    rays = (np.linalg.inv(K) @ np.column_stack([kps["x"], kps["y"], np.ones(len(kps))]).T).T @ R   # rows: R^T d
    lam = -C[2] / rays[:, 2]
    X = C + lam[:, None] * rays
    dist = np.linalg.norm(X - C, axis=1)
    sf = np.array([np.float32(SCALE) ** o for o in range(N_LEVELS)], np.float32)
    dmax = dist * sf[kps["octave"]]
    n = (C - X) / dist[:, None]
    return dict(mp_id=np.arange(len(kps), dtype=np.uint32), mp_pos=X.astype(np.float32), mp_normal=n.astype(np.float32),
                mp_min_dist=(dmax / sf[-1]).astype(np.float32), mp_max_dist=dmax.astype(np.float32), mp_desc=desc.copy(),
                scale_factors=sf, fx=float(np.float32(F)), fy=float(np.float32(F)), cx=float(np.float32(CX)), cy=float(np.float32(CY)),
                min_xy=np.array([0, 0], np.float32), max_xy=np.array([W, H], np.float32))


def track(frames, T0, extract, match_projected, pose_only, min_desc_dist=60.0, max_reproj=15.0):
    """returns (poses (n,4,4) f32 frame<-world incl. the given first one, per-frame match / inlier counts)"""
    kps, desc = extract(frames[0])
    scene = make_map(kps, desc, T0)
    poses, stats = [T0.astype(np.float32)], []
    pose = T0.astype(np.float32)
    for img in frames[1:]:
        kps, desc = extract(img)
        sc = dict(scene, kp_xy=np.stack([kps["x"], kps["y"]], 1).astype(np.float32), kp_octave=kps["octave"].astype(np.int32), kp_desc=desc,
                  pose44=pose.reshape(16))
        m, _ = match_projected(sc, min_desc_dist, max_reproj)
        q, tr = m["queryIdx"], m["trainIdx"]
        pb = dict(pose44=pose.reshape(16).copy(), points3=scene["mp_pos"][tr], obs_uv=sc["kp_xy"][q], obs_ur=np.zeros(len(m), np.float32),
                  obs_stereo=np.zeros(len(m), np.uint8), obs_inv_sigma2=(np.float32(1) / scene["scale_factors"][sc["kp_octave"][q]]).astype(np.float32),
                  stable=np.ones(len(m), np.uint8), fx=scene["fx"], fy=scene["fy"], cx=scene["cx"], cy=scene["cy"], bf=float(np.float32(0.12 * F)),
                  marker_pose44=np.zeros((0, 16), np.float32), marker_size=np.zeros(0, np.float32), marker_corners=np.zeros((0, 8), np.float32))
        r = pose_only(pb)
        pose = np.asarray(r["pose44"], np.float32).reshape(4, 4)
        poses.append(pose)
        stats.append((len(m), int(r["n_good"])))
    return np.array(poses), stats


def track_full(frames, T0, extract, track_frame):
    """The tracker's full per-frame sequence (System::_11166622111371682966, src/utils/system.cpp:6460-6960) over a clip: the map is made
    from the first frame; every following frame is extracted and handed to `track_frame(scene) -> dict(pose44, matches, n_good, ..)`
    (search by projection from the previous frame -> solvePnp -> local-map search -> solvePnp) together with the previous frame's
    keypoints and their map-point assignment, which the matches of the previous call define (Frame::ids, :6950-6956).
    Returns (poses (n,4,4) f32 incl. the given first one, per-frame (n_tbp, n_matches, n_good))."""
    kps, desc = extract(frames[0])
    scene = make_map(kps, desc, T0)
    scene["mp_stable"] = np.ones(len(kps), np.uint8)
    scene["mp_local"] = np.ones(len(kps), np.uint8)
    scene["bf"] = 0.0
    prev = dict(prev_octave=kps["octave"].astype(np.int32), prev_desc=desc, prev_mp_row=np.arange(len(kps), dtype=np.int32))
    pose = T0.astype(np.float32)
    poses, stats = [pose], []
    for img in frames[1:]:
        kps, desc = extract(img)
        sc = dict(scene, kp_xy=np.stack([kps["x"], kps["y"]], 1).astype(np.float32), kp_octave=kps["octave"].astype(np.int32), kp_desc=desc,
                  pose44=pose.reshape(16), **prev)
        r = track_frame(sc)
        pose = np.asarray(r["pose44"], np.float32).reshape(4, 4)
        poses.append(pose)
        m = r["matches"]
        rows = np.full(len(kps), -1, np.int32)
        rows[m["queryIdx"]] = m["trainIdx"]          # mp_id == row in make_map
        prev = dict(prev_octave=kps["octave"].astype(np.int32), prev_desc=desc, prev_mp_row=rows)
        stats.append((int(r["n_tbp"]), len(m), int(r["n_good"])))
    return np.array(poses), stats


def centres(poses):
    return np.array([-(T[:3, :3].astype(np.float64).T @ T[:3, 3].astype(np.float64)) for T in poses])


def ate(poses, gt):
    """absolute trajectory error: RMSE of the camera centres against ground truth (same world frame, no alignment needed)"""
    d = centres(poses) - centres(gt)
    return float(np.sqrt((d ** 2).sum(1).mean()))
