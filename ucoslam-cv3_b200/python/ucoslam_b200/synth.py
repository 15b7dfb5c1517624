"""Seeded synthetic workloads shared by bench.py and the tests: no oracle or reference code here."""
import numpy as np

def _rodrigues(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def synth_ba_problem(seed, n_poses=12, n_fixed=2, n_points=2000, stereo_frac=0.0, outlier_frac=0.02, px_sigma=0.5,
                     pose_noise=(0.01, 0.5), point_noise=0.02, w=640, h=480, f=525.0, bf=0.12 * 525.0, max_obs=None):
    """A seeded local-BA window: cameras on an arc looking at a point cloud; mono (and optionally stereo)
    observations with pixel noise scaled by pyramid octave, a few gross outliers, perturbed initial poses/points.
    All inputs are stored in f32 exactly as the reference's data model holds them (pose_f2g CV_32F, Point3f, KeyPoint.pt)."""
    rng = np.random.default_rng(seed)
    fx = fy = np.float32(f)
    cx, cy = np.float32(w / 2 - 0.5), np.float32(h / 2 - 0.5)
    poses_gt = []
    for i in range(n_poses):
        a = (i - n_poses / 2) * 0.06
        R = _rodrigues(np.array([0.02 * np.sin(i), a, 0.01 * i]))
        C = np.array([1.5 * np.sin(a * 2), 0.05 * np.cos(i), -0.3 * np.cos(a)])
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = -R @ C
        poses_gt.append(T)
    poses_gt = np.array(poses_gt)
    pts_gt = np.c_[rng.uniform(-3, 3, n_points), rng.uniform(-2, 2, n_points), rng.uniform(2.5, 8, n_points)]
    obs_pose, obs_point, obs_uv, obs_ur, obs_st, obs_inv = [], [], [], [], [], []
    for p in range(n_points):
        cnt = 0
        for i in range(n_poses):
            Xc = poses_gt[i][:3, :3] @ pts_gt[p] + poses_gt[i][:3, 3]
            if Xc[2] < 0.5:
                continue
            u, v = f * Xc[0] / Xc[2] + cx, f * Xc[1] / Xc[2] + cy
            if not (20 < u < w - 20 and 20 < v < h - 20) or rng.random() < 0.25:
                continue
            octave = int(rng.integers(0, 8))
            s = 1.2 ** octave
            nu, nv = rng.normal(0, px_sigma * s, 2)
            if rng.random() < outlier_frac:
                nu, nv = rng.uniform(-40, 40, 2)
            st = rng.random() < stereo_frac
            obs_pose.append(i); obs_point.append(p); obs_uv.append((u + nu, v + nv))
            obs_ur.append(u + nu - bf / Xc[2] + rng.normal(0, px_sigma * s) if st else 0.0)
            obs_st.append(1 if st else 0)
            obs_inv.append(np.float32(1.0) / np.float32(np.float32(1.2) ** octave))
            cnt += 1
            if max_obs and cnt >= max_obs:
                break
    poses0 = poses_gt.copy()
    for i in range(n_fixed, n_poses):
        dR = _rodrigues(rng.normal(0, np.deg2rad(pose_noise[1]), 3))
        poses0[i][:3, :3] = dR @ poses0[i][:3, :3]
        poses0[i][:3, 3] += rng.normal(0, pose_noise[0], 3)
    pts0 = pts_gt + rng.normal(0, point_noise, pts_gt.shape)
    fixed = np.zeros(n_poses, np.uint8)
    fixed[:n_fixed] = 1
    return dict(poses44=poses0.reshape(n_poses, 16).astype(np.float32), fixed=fixed, points3=pts0.astype(np.float32),
                obs_pose=np.array(obs_pose, np.int32), obs_point=np.array(obs_point, np.int32),
                obs_uv=np.array(obs_uv, np.float32).reshape(-1, 2), obs_ur=np.array(obs_ur, np.float32),
                obs_stereo=np.array(obs_st, np.uint8), obs_inv_sigma2=np.array(obs_inv, np.float32),
                fx=float(fx), fy=float(fy), cx=float(cx), cy=float(cy), bf=float(np.float32(bf)),
                poses_gt=poses_gt, points_gt=pts_gt)


def synth_pnp_problem(seed, n_matches=800, stereo_frac=0.0, outlier_frac=0.1, unstable_frac=0.3, n_markers=0, px_sigma=0.5,
                      pose_noise=(0.03, 1.5), w=640, h=480, f=525.0, bf=0.12 * 525.0):
    """A seeded pose-only problem as the tracker hands it to PnPSolver::solvePnp: one camera, n_matches
    (keypoint, map point) pairs with octave-scaled pixel noise and gross outliers (wrong associations), optionally stereo
    observations and ArUco markers with known map pose.  f32 containers as in the reference's data model."""
    rng = np.random.default_rng(seed)
    fx = fy = np.float32(f)
    cx, cy = np.float32(w / 2 - 0.5), np.float32(h / 2 - 0.5)
    T = np.eye(4)
    T[:3, :3] = _rodrigues(rng.normal(0, 0.2, 3))
    T[:3, 3] = rng.normal(0, 0.5, 3)
    pts, uv, ur, st, inv, stable = [], [], [], [], [], []
    while len(pts) < n_matches:
        u, v, z = rng.uniform(20, w - 20), rng.uniform(20, h - 20), rng.uniform(1.0, 8.0)
        Xc = np.array([(u - cx) / f * z, (v - cy) / f * z, z])
        X = T[:3, :3].T @ (Xc - T[:3, 3])
        octave = int(rng.integers(0, 8))
        s = 1.2 ** octave
        nu, nv = rng.normal(0, px_sigma * s, 2)
        if rng.random() < outlier_frac:
            nu, nv = rng.uniform(-60, 60, 2)
        is_st = rng.random() < stereo_frac
        pts.append(X); uv.append((u + nu, v + nv))
        depth = np.float32(z + (rng.normal(0, 0.01 * z) if is_st else 0))
        ur.append(np.float32(u + nu) - np.float32(bf) / depth if is_st else 0.0)   # kp_ur = kpt.pt.x - mbf/depth, pnpsolver.cpp:226
        st.append(1 if is_st else 0)
        inv.append(np.float32(1.0) / np.float32(np.float32(1.2) ** octave))
        stable.append(0 if rng.random() < unstable_frac else 1)
    m_pose, m_size, m_corners = [], [], []
    for m in range(n_markers):
        size = np.float32(rng.uniform(0.1, 0.3))
        Mc = np.eye(4)   # marker -> camera
        Mc[:3, :3] = _rodrigues(rng.normal(0, 0.3, 3))
        Mc[:3, 3] = [rng.uniform(-0.8, 0.8), rng.uniform(-0.5, 0.5), rng.uniform(1.0, 3.0)]
        g2m = np.linalg.inv(T) @ Mc
        hs = size / 2
        corners = []
        for c in ((-hs, hs, 0), (hs, hs, 0), (hs, -hs, 0), (-hs, -hs, 0)):
            Xc = Mc[:3, :3] @ np.array(c) + Mc[:3, 3]
            corners += [f * Xc[0] / Xc[2] + cx + rng.normal(0, 0.3), f * Xc[1] / Xc[2] + cy + rng.normal(0, 0.3)]
        m_pose.append(g2m.reshape(16)); m_size.append(size); m_corners.append(corners)
    T0 = T.copy()
    T0[:3, :3] = _rodrigues(rng.normal(0, np.deg2rad(pose_noise[1]), 3)) @ T0[:3, :3]
    T0[:3, 3] += rng.normal(0, pose_noise[0], 3)
    return dict(pose44=T0.reshape(16).astype(np.float32), points3=np.array(pts, np.float32).reshape(-1, 3),
                obs_uv=np.array(uv, np.float32).reshape(-1, 2), obs_ur=np.array(ur, np.float32),
                obs_stereo=np.array(st, np.uint8), obs_inv_sigma2=np.array(inv, np.float32), stable=np.array(stable, np.uint8),
                fx=float(fx), fy=float(fy), cx=float(cx), cy=float(cy), bf=float(np.float32(bf)),
                marker_pose44=np.array(m_pose, np.float32).reshape(-1, 16), marker_size=np.array(m_size, np.float32),
                marker_corners=np.array(m_corners, np.float32).reshape(-1, 8), pose_gt=T)
