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


def add_markers(pb, seed, n_markers=4, size=0.25, corner_sigma=0.3, pose_noise=(0.02, 1.0), opt_weight=0.5, min_markers=5, w=640, h=480, coplanar=False):
    """ArUco markers for a BA problem made by synth_ba_problem / synth_global_ba: marker poses (global <- marker) in front of the
    cameras, an observation (4 undistorted corners) in every keyframe that sees all four, perturbed initial marker poses, and the
    per-observation weight the reference derives per keyframe (globaloptimizer_g2o.cpp:276-297: markersOptWeight share of the frame's
    keypoint weight spread over its markers' 8 residuals).  Adds marker_pose44, marker_size, mobs_marker / _pose / _corners / _weight."""
    rng = np.random.default_rng(seed)
    T = pb["poses_gt"]
    P = len(T)
    f, cx, cy = pb["fx"], pb["cx"], pb["cy"]
    hs = size / 2
    local = np.array([[-hs, hs, 0], [hs, hs, 0], [hs, -hs, 0], [-hs, -hs, 0]])
    g2m_gt, mm, mp, mc = [], [], [], []
    for m in range(n_markers):
        cam = T[int(rng.integers(0, P))]
        Mc = np.eye(4)   # marker -> that camera
        Mc[:3, :3] = _rodrigues(rng.normal(0, 0.3, 3))
        Mc[:3, 3] = [rng.uniform(-0.6, 0.6), rng.uniform(-0.4, 0.4), rng.uniform(1.5, 3.5)]
        G = np.linalg.inv(cam) @ Mc
        if coplanar and g2m_gt:      # the InPlaneMarkers scene: every marker lies in the first one's plane with the same normal (in-plane shift + turn)
            S = np.eye(4)
            S[:3, :3] = _rodrigues(np.array([0, 0, rng.uniform(-0.5, 0.5)]))
            S[:3, 3] = [rng.uniform(-1.2, 1.2), rng.uniform(-0.8, 0.8), 0.0]
            G = g2m_gt[0] @ S
        g2m_gt.append(G)
        for i in range(P):
            Xc = (T[i] @ G @ np.c_[local, np.ones(4)].T).T[:, :3]
            if (Xc[:, 2] < 0.3).any():
                continue
            uv = np.c_[f * Xc[:, 0] / Xc[:, 2] + cx, f * Xc[:, 1] / Xc[:, 2] + cy]
            if ((uv[:, 0] < 5) | (uv[:, 0] > w - 5) | (uv[:, 1] < 5) | (uv[:, 1] > h - 5)).any():
                continue
            mm.append(m); mp.append(i); mc.append((uv + rng.normal(0, corner_sigma, (4, 2))).reshape(8))
    mm, mp = np.array(mm, np.int32), np.array(mp, np.int32)
    kpw = np.zeros(P)
    np.add.at(kpw, pb["obs_pose"], np.where(pb["obs_stereo"] != 0, 3.0, 2.0) * pb["obs_inv_sigma2"].astype(np.float64))
    n_in_frame = np.bincount(mp, minlength=P)
    wgt = np.ones(len(mm))
    for k in range(len(mm)):
        fr = mp[k]
        if kpw[fr] > 40 and n_in_frame[fr] > 0:
            wgt[k] = opt_weight * min(1.0, n_in_frame[fr] / min_markers) * kpw[fr] / (n_in_frame[fr] * 8)
    g0 = []
    for G in g2m_gt:
        G0 = G.copy()
        G0[:3, :3] = _rodrigues(rng.normal(0, np.deg2rad(pose_noise[1]), 3)) @ G0[:3, :3]
        G0[:3, 3] += rng.normal(0, pose_noise[0], 3)
        g0.append(G0.reshape(16))
    out = dict(pb)
    out.update(marker_pose44=np.array(g0, np.float32).reshape(-1, 16), marker_size=np.full(n_markers, size, np.float32), mobs_marker=mm, mobs_pose=mp,
               mobs_corners=np.array(mc, np.float32).reshape(-1, 8), mobs_weight=wgt.astype(np.float32), marker_gt=np.array(g2m_gt))
    return out


def mix_cameras(pb, seed, frac=0.5, scale=(1.6, 1.45), shift=(37.0, -21.0), bl_scale=1.3):
    """Turn a single-camera BA problem (synth_ba_problem / synth_global_ba, optionally with add_markers) into a window whose keyframes were
    taken with two cameras: about `frac` of the keyframes get a second camera (fx, fy scaled, principal point shifted, another stereo
    baseline) and their observations are re-projected through it (same rays, the pixel noise scales with the focal length).  Adds pose_cam
    (n_poses x 5: fx fy cx cy bf per keyframe) - the table uco_ba_problem::pose_cam / GlobalOptimizerG2O's per-edge ImageParams stand for."""
    rng = np.random.default_rng(seed)
    P = len(pb["fixed"])
    second = rng.random(P) < frac
    second[0], second[-1] = False, True                      # both cameras are present
    c1 = np.array([pb["fx"], pb["fy"], pb["cx"], pb["cy"], pb["bf"]], np.float64)
    c2 = np.array([c1[0] * scale[0], c1[1] * scale[1], c1[2] + shift[0], c1[3] + shift[1], c1[4] * scale[0] * bl_scale])
    out = dict(pb)
    cam = np.where(second[:, None], c2[None, :], c1[None, :])
    uv = np.array(pb["obs_uv"], np.float64).reshape(-1, 2)
    ur = np.array(pb["obs_ur"], np.float64).reshape(-1)
    st = np.asarray(pb["obs_stereo"]).astype(bool)
    sel = second[np.asarray(pb["obs_pose"])]
    disp = uv[:, 0] - ur                                      # bf / depth
    uv2 = uv.copy()
    uv2[sel, 0] = (uv[sel, 0] - c1[2]) / c1[0] * c2[0] + c2[2]
    uv2[sel, 1] = (uv[sel, 1] - c1[3]) / c1[1] * c2[1] + c2[3]
    ur2 = ur.copy()
    both = sel & st
    ur2[both] = uv2[both, 0] - disp[both] / c1[4] * c2[4] if c1[4] else ur[both]
    out.update(obs_uv=uv2.astype(np.float32), obs_ur=ur2.astype(np.float32), pose_cam=cam.astype(np.float32))
    if len(pb.get("marker_size", ())):
        mc = np.array(pb["mobs_corners"], np.float64).reshape(-1, 4, 2)
        msel = second[np.asarray(pb["mobs_pose"])]
        mc[msel, :, 0] = (mc[msel, :, 0] - c1[2]) / c1[0] * c2[0] + c2[2]
        mc[msel, :, 1] = (mc[msel, :, 1] - c1[3]) / c1[1] * c2[1] + c2[3]
        out["mobs_corners"] = mc.reshape(-1, 8).astype(np.float32)
    return out


def add_plane_edges(pb, ref_in_window=True):
    """The InPlaneMarkers option for a problem made by add_markers(coplanar=True): the reference marker is the one with most observations
    (globaloptimizer_g2o.cpp:362-368); every other marker gets one planar edge of weight 0.33 * (sum of marker weights x 8 + sum of the
    keyframes' keypoint weights) / (4 * (markers - 1)) (:381-382).  ref_in_window=False takes the reference OUT of the window (its
    vertex, observations and weights go; its pose stays as the fixed plane_ref_pose44) - the case of :371-379."""
    out = dict(pb)
    nm = len(pb["marker_size"])
    counts = np.bincount(pb["mobs_marker"], minlength=nm)
    ref = int(np.argmax(counts))
    kpw = np.zeros(len(pb["fixed"]))
    np.add.at(kpw, pb["obs_pose"], np.where(np.asarray(pb["obs_stereo"]) != 0, 3.0, 2.0) * np.asarray(pb["obs_inv_sigma2"], np.float64))
    if not ref_in_window:
        keep = np.asarray(pb["mobs_marker"]) != ref
        remap = np.cumsum(np.arange(nm) != ref) - 1
        out.update(plane_ref_pose44=np.asarray(pb["marker_gt"][ref], np.float32).reshape(16), marker_pose44=np.delete(pb["marker_pose44"], ref, 0),
                   marker_size=np.delete(pb["marker_size"], ref, 0), mobs_marker=remap[np.asarray(pb["mobs_marker"])[keep]].astype(np.int32),
                   mobs_pose=np.asarray(pb["mobs_pose"])[keep], mobs_corners=np.asarray(pb["mobs_corners"])[keep], mobs_weight=np.asarray(pb["mobs_weight"])[keep],
                   marker_gt=np.delete(pb["marker_gt"], ref, 0), plane_ref=-1, plane_other=np.arange(nm - 1, dtype=np.int32))
    else:
        out.update(plane_ref=ref, plane_other=np.array([m for m in range(nm) if m != ref], np.int32))
    total = 8.0 * np.asarray(out["mobs_weight"], np.float64).sum() + kpw.sum()
    n_info = len(out["marker_size"]) + (0 if ref_in_window else 1)          # marker_info.size(): the window's markers + the added reference
    out["plane_weight"] = 0.33 * total / (4.0 * (n_info - 1))
    return out


def synth_global_ba(seed, n_kf=500, n_points=50000, obs_per_point=6, n_fixed=2, loop_m=50.0, px_sigma=0.5, pose_noise=(0.01, 0.5),
                    point_noise=0.02, outlier_frac=0.01, w=640, h=480, f=525.0):
    """BASELINE config 5: n_kf keyframes on a closed loop of loop_m metres looking outwards, n_points landmarks each
    seen by obs_per_point consecutive keyframes (the reduced camera system is block-banded plus the loop-closing corner), pixel
    noise px_sigma, perturbed initial poses / points, the first n_fixed keyframes fixed.  Vectorised; f32 containers."""
    rng = np.random.default_rng(seed)
    cx, cy = np.float32(w / 2 - 0.5), np.float32(h / 2 - 0.5)
    r = loop_m / (2 * np.pi)
    th = 2 * np.pi * np.arange(n_kf) / n_kf
    zc = np.stack([np.cos(th), np.zeros(n_kf), np.sin(th)], 1)
    xc = np.stack([np.sin(th), np.zeros(n_kf), -np.cos(th)], 1)
    yc = np.tile(np.array([0.0, 1.0, 0.0]), (n_kf, 1))
    C = r * zc
    R = np.stack([xc, yc, zc], 1)                      # (n_kf, 3, 3) rows = camera axes in the world
    t = -np.einsum("kij,kj->ki", R, C)
    base = rng.integers(0, n_kf, n_points)
    thm = 2 * np.pi * (base + (obs_per_point - 1) / 2) / n_kf
    zm = np.stack([np.cos(thm), np.zeros(n_points), np.sin(thm)], 1)
    xm = np.stack([np.sin(thm), np.zeros(n_points), -np.cos(thm)], 1)
    depth = rng.uniform(3, 8, n_points)
    pts_gt = r * zm + depth[:, None] * zm + rng.uniform(-1.5, 1.5, n_points)[:, None] * xm
    pts_gt[:, 1] += rng.uniform(-1.5, 1.5, n_points)
    op, ol, ouv, oinv = [], [], [], []
    for k in range(obs_per_point):
        kf = (base + k) % n_kf
        Xc = np.einsum("nij,nj->ni", R[kf], pts_gt) + t[kf]
        u = f * Xc[:, 0] / Xc[:, 2] + cx
        v = f * Xc[:, 1] / Xc[:, 2] + cy
        octave = rng.integers(0, 8, n_points)
        sc = 1.2 ** octave
        noise = rng.normal(0, px_sigma, (n_points, 2)) * sc[:, None]
        out = rng.random(n_points) < outlier_frac
        noise[out] = rng.uniform(-40, 40, (int(out.sum()), 2))
        ok = (Xc[:, 2] > 0.5) & (u > 20) & (u < w - 20) & (v > 20) & (v < h - 20)
        op.append(kf[ok]); ol.append(np.nonzero(ok)[0]); ouv.append(np.stack([u + noise[:, 0], v + noise[:, 1]], 1)[ok])
        oinv.append((np.float32(1.0) / (np.float32(1.2) ** octave.astype(np.float32)))[ok])
    obs_pose = np.concatenate(op).astype(np.int32)
    obs_point = np.concatenate(ol).astype(np.int32)
    perm = np.argsort(obs_point * n_kf + obs_pose, kind="stable")   # landmark-major like a map walk; any order is accepted
    poses_gt = np.tile(np.eye(4), (n_kf, 1, 1))
    poses_gt[:, :3, :3] = R
    poses_gt[:, :3, 3] = t
    poses0 = poses_gt.copy()
    for i in range(n_fixed, n_kf):
        dR = _rodrigues(rng.normal(0, np.deg2rad(pose_noise[1]), 3))
        poses0[i][:3, :3] = dR @ poses0[i][:3, :3]
        poses0[i][:3, 3] += rng.normal(0, pose_noise[0], 3)
    pts0 = pts_gt + rng.normal(0, point_noise, pts_gt.shape)
    fixed = np.zeros(n_kf, np.uint8)
    fixed[:n_fixed] = 1
    M = len(obs_pose)
    return dict(poses44=poses0.reshape(n_kf, 16).astype(np.float32), fixed=fixed, points3=pts0.astype(np.float32),
                obs_pose=obs_pose[perm], obs_point=obs_point[perm], obs_uv=np.concatenate(ouv).astype(np.float32)[perm],
                obs_ur=np.zeros(M, np.float32), obs_stereo=np.zeros(M, np.uint8),
                obs_inv_sigma2=np.concatenate(oinv).astype(np.float32)[perm],
                fx=float(np.float32(f)), fy=float(np.float32(f)), cx=float(cx), cy=float(cy), bf=float(np.float32(0.12 * f)),
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


def synth_projection_scene(seed, n_kp=2000, n_mp=3000, n_levels=8, scale=1.2, w=640, h=480, f=525.0, dup_frac=0.1, clutter=0.3):
    """A frame (undistorted keypoints with octaves and 256-bit descriptors) and the local map handed to
    Map::matchFrameToMapPoints: most map points project next to a keypoint of a compatible octave and carry
    that keypoint's descriptor with a few flipped bits; some share their keypoint with another map point (filter_ambiguous_query),
    some look away, lie behind the camera, outside the image or outside their scale-invariance range; keypoints cluster so that
    several candidates fall inside a search radius (best / second-best bookkeeping)."""
    rng = np.random.default_rng(seed)
    cx, cy = np.float32(w / 2 - 0.5), np.float32(h / 2 - 0.5)
    sf = np.array([np.float32(scale) ** i for i in range(n_levels)], np.float32)
    n_clusters = max(1, int(n_kp * clutter / 4))
    kxy = rng.uniform([20, 20], [w - 20, h - 20], (n_kp, 2))
    own = rng.integers(0, n_kp - n_clusters * 4, n_clusters)
    for c in range(n_clusters):   # small clumps of keypoints around some others
        kxy[n_kp - 4 * c - 4:n_kp - 4 * c] = kxy[own[c]] + rng.normal(0, 2.5, (4, 2))
    kxy = kxy.astype(np.float32)
    koct = rng.integers(0, n_levels, n_kp).astype(np.int32)
    kdesc = rng.integers(0, 256, (n_kp, 32), dtype=np.uint8)
    T = np.eye(4)
    T[:3, :3] = _rodrigues(rng.normal(0, 0.3, 3))
    T[:3, 3] = rng.normal(0, 1.0, 3)
    pose = T.astype(np.float32)
    Rm, t = pose[:3, :3].astype(np.float64), pose[:3, 3].astype(np.float64)
    C = -Rm.T @ t
    pos, nrm, dmin, dmax, mdesc = [], [], [], [], []
    for i in range(n_mp):
        kind = rng.random()
        kp = int(rng.integers(0, n_kp)) if (i == 0 or rng.random() > dup_frac) else last_kp
        last_kp = kp
        z = rng.uniform(1.0, 8.0)
        u, v = kxy[kp] + rng.normal(0, 1.2, 2)
        Xc = np.array([(u - cx) / f * z, (v - cy) / f * z, z])
        if kind < 0.05:
            Xc[2] = -Xc[2]                                   # behind the camera
        elif kind < 0.10:
            Xc[0] += 3 * z                                   # projects outside the image
        X = Rm.T @ (Xc - t)
        dist = np.linalg.norm(Xc)
        n = (C - X) / max(np.linalg.norm(C - X), 1e-9)
        if kind < 0.18:
            n = _rodrigues(rng.normal(0, 1.0, 3)) @ n        # looks (more or less) away
        else:
            n = _rodrigues(rng.normal(0, 0.15, 3)) @ n
        octv = int(koct[kp])
        mx = dist * float(sf[octv]) * rng.uniform(0.92, 1.0)  # predictScale() lands on the keypoint's octave (or the one above)
        mn = mx / float(sf[-1]) * rng.uniform(0.7, 1.0)
        if 0.18 <= kind < 0.22:
            mx *= 0.5                                        # too far for its invariance range
        d = kdesc[kp].copy()
        for b in rng.integers(0, 256, int(rng.integers(0, 70))):
            d[b >> 3] ^= 1 << (b & 7)
        pos.append(X); nrm.append(n); dmin.append(mn); dmax.append(mx); mdesc.append(d)
    return dict(kp_xy=kxy, kp_octave=koct, kp_desc=kdesc, mp_id=(np.arange(n_mp) * 3 + 7).astype(np.uint32),
                mp_pos=np.array(pos, np.float32).reshape(-1, 3), mp_normal=np.array(nrm, np.float32).reshape(-1, 3),
                mp_min_dist=np.array(dmin, np.float32), mp_max_dist=np.array(dmax, np.float32),
                mp_desc=np.array(mdesc, np.uint8).reshape(-1, 32), scale_factors=sf, pose44=pose.reshape(16),
                fx=float(np.float32(f)), fy=float(np.float32(f)), cx=float(cx), cy=float(cy),
                min_xy=np.array([0, 0], np.float32), max_xy=np.array([w, h], np.float32))


def synth_track_scene(seed, n_kp=2000, n_mp=3000, n_prev=1800, prior_noise=(0.004, 0.01), **kw):
    """A tracking problem as the tracker's main branch sees it (System::_11166622111371682966, src/utils/system.cpp:6460-6960): the
    projection scene of synth_projection_scene (current frame + map, consistent with a true pose up to ~1 px) plus a PREVIOUS frame
    whose keypoints carry map points (prev_mp_row, some -1, a few shared), descriptors a few bits off the map point's and mostly
    the octave of the keypoint the point lands on, and a pose prior a few pixels off the true pose."""
    sc = synth_projection_scene(seed, n_kp=n_kp, n_mp=n_mp, **kw)
    rng = np.random.default_rng(seed + 77)
    n_mp = len(sc["mp_id"])
    T = sc["pose44"].reshape(4, 4).astype(np.float64)
    # which current keypoint each map point lands next to (nearest projection), to give the previous keypoint a plausible octave
    X = sc["mp_pos"].astype(np.float64)
    Xc = X @ T[:3, :3].T + T[:3, 3]
    z = np.where(np.abs(Xc[:, 2]) < 1e-9, 1e-9, Xc[:, 2])
    uv = np.stack([Xc[:, 0] / z * sc["fx"] + sc["cx"], Xc[:, 1] / z * sc["fy"] + sc["cy"]], 1)
    kxy = sc["kp_xy"].astype(np.float64)
    rows = rng.integers(0, n_mp, n_prev).astype(np.int32)
    rows[rng.random(n_prev) < 0.15] = -1
    poct = np.zeros(n_prev, np.int32)
    pdesc = np.zeros((n_prev, 32), np.uint8)
    n_levels = len(sc["scale_factors"])
    for i in range(n_prev):
        r = rows[i]
        if r < 0:
            poct[i] = rng.integers(0, n_levels)
            pdesc[i] = rng.integers(0, 256, 32)
            continue
        j = int(np.argmin(((kxy - uv[r]) ** 2).sum(1)))
        poct[i] = sc["kp_octave"][j] if rng.random() > 0.08 else rng.integers(0, n_levels)
        d = sc["kp_desc"][j].copy() if rng.random() > 0.2 else sc["mp_desc"][r].copy()
        for b in rng.integers(0, 256, int(rng.integers(0, 50))):
            d[b >> 3] ^= 1 << (b & 7)
        pdesc[i] = d
    dT = np.eye(4)
    dT[:3, :3] = _rodrigues(rng.normal(0, prior_noise[0], 3))
    dT[:3, 3] = rng.normal(0, prior_noise[1], 3)
    sc["pose_true44"] = sc["pose44"].copy()
    sc["pose44"] = (dT @ T).astype(np.float32).reshape(16)
    sc.update(prev_octave=poct, prev_desc=pdesc, prev_mp_row=rows, mp_stable=(rng.random(n_mp) > 0.2).astype(np.uint8),
              mp_local=(rng.random(n_mp) > 0.1).astype(np.uint8), bf=0.0)
    return sc


def synth_stereo(seed, w=640, h=480, n=1500, max_disp=48.0, outlier_frac=0.2, tie_frac=0.05):
    """A seeded rectified stereo pair as FrameExtractor::processStereo sees it: block-noise images where the
    right image is the left one shifted by a smoothly varying sub-pixel disparity, left keypoints with octaves and 256-bit
    descriptors, and right detections = most left keypoints moved by the disparity (pixel noise, octave changes, descriptor bit
    flips) plus unrelated ones; some keypoints sit exactly on .5 rows / columns and some right descriptors are duplicated so that
    distance ties and rounding edges are exercised.  Returns dict(img_l, img_r, kps_l, desc_l, kps_r, desc_r, bl, fx)."""
    from . import KP_DTYPE
    rng = np.random.default_rng(seed)
    margin = max(64, int(max_disp) + 2)
    coarse = np.kron(rng.integers(0, 256, (h // 8 + 2, (w + margin) // 8 + 2)), np.ones((8, 8)))[:h, :w + margin]
    left_wide = np.clip(0.6 * coarse + 0.4 * rng.integers(0, 256, (h, w + margin)), 0, 255)
    img_l = left_wide[:, :w].astype(np.uint8)
    disp_row = 8.0 + (max_disp - 8.0) * (0.5 + 0.5 * np.sin(np.arange(h) / h * 2 * np.pi))   # disparity per row
    img_r = np.empty((h, w), np.uint8)
    xs = np.arange(w)
    for y in range(h):
        src = xs + disp_row[y]                 # right(x) = left(x + d)
        x0 = np.floor(src).astype(int); a = src - x0
        img_r[y] = np.clip((1 - a) * left_wide[y, x0] + a * left_wide[y, x0 + 1] + 0.5, 0, 255).astype(np.uint8)
    kps_l = np.zeros(n, KP_DTYPE)
    kps_l["x"] = rng.uniform(16, w - 17, n).astype(np.float32)
    kps_l["y"] = rng.uniform(16, h - 17, n).astype(np.float32)
    half = rng.random(n) < 0.1
    kps_l["y"][half] = np.floor(kps_l["y"][half]) + 0.5
    kps_l["x"][half[::-1]] = np.floor(kps_l["x"][half[::-1]]) + 0.5
    kps_l["octave"] = rng.integers(0, 8, n)
    kps_l["size"] = 31.0; kps_l["angle"] = rng.uniform(0, 360, n); kps_l["response"] = rng.integers(20, 200, n); kps_l["class_id"] = -1
    desc_l = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    keep = np.nonzero(rng.random(n) > outlier_frac)[0]
    m = len(keep)
    kps_r = np.zeros(m, KP_DTYPE)
    yl = kps_l["y"][keep]
    d = disp_row[np.clip(np.round(yl).astype(int), 0, h - 1)]
    xerr = np.where(rng.random(m) < 0.15, rng.uniform(-10, 10, m), rng.uniform(-1.5, 1.5, m))   # some minima fall on the search border
    kps_r["x"] = (kps_l["x"][keep] - d + xerr).astype(np.float32)
    kps_r["y"] = (yl + rng.choice([0.0, 0.0, 0.0, 0.3, -0.3, 0.6, -0.6], m)).astype(np.float32)
    kps_r["octave"] = kps_l["octave"][keep] + rng.choice([0, 0, 0, 1, -1, 2], m)
    kps_r["size"] = 31.0; kps_r["angle"] = kps_l["angle"][keep]; kps_r["response"] = kps_l["response"][keep]; kps_r["class_id"] = -1
    bits = np.unpackbits(desc_l[keep], axis=1)
    for r in range(m):
        nf = int(rng.integers(0, 70))
        if nf:
            bits[r, rng.choice(256, nf, replace=False)] ^= 1
    desc_r = np.packbits(bits, axis=1)
    # unrelated right detections + duplicated rows (ties) appended, then everything shuffled
    extra = int(n * outlier_frac)
    kx = np.zeros(extra, KP_DTYPE)
    kx["x"] = rng.uniform(16, w - 17, extra).astype(np.float32); kx["y"] = rng.uniform(16, h - 17, extra).astype(np.float32)
    kx["octave"] = rng.integers(0, 8, extra); kx["size"] = 31.0; kx["class_id"] = -1
    dx = rng.integers(0, 256, (extra, 32), dtype=np.uint8)
    nt = int(m * tie_frac)
    dup = rng.integers(0, m, nt)
    kt = kps_r[dup].copy()
    kt["x"] -= rng.uniform(0.0, 3.0, nt).astype(np.float32)
    kps_r = np.concatenate([kps_r, kx, kt]); desc_r = np.concatenate([desc_r, dx, desc_r[dup]])
    ok = (kps_r["x"] >= 16) & (kps_r["x"] <= w - 17)      # ORB keypoints stay 16 px inside the image
    kps_r, desc_r = kps_r[ok], desc_r[ok]
    perm = rng.permutation(len(kps_r))
    return dict(img_l=np.ascontiguousarray(img_l), img_r=np.ascontiguousarray(img_r), kps_l=kps_l, desc_l=np.ascontiguousarray(desc_l),
                kps_r=np.ascontiguousarray(kps_r[perm]), desc_r=np.ascontiguousarray(desc_r[perm]), bl=np.float32(0.12), fx=np.float32(525.0))


def synth_two_view(seed, n=1500, w=640, h=480, f=525.0, baseline=0.25, px_sigma=0.6, outlier_frac=0.15, far_frac=0.15):
    """Two keyframes and their matches as the mapper hands them to ucoslam::Triangulate: 3-D points in front of
    camera 1 (some far away: parallax below the reference's gate), camera 2 displaced by `baseline` and slightly rotated, pixel noise
    scaled by octave, a share of wrong matches.  Returns dict(kps_train, kps_query, matches, K_train, K_query, RT, sf_train, sf_query,
    xyz_gt (n,3) with NaN for wrong matches)."""
    from . import KP_DTYPE, MATCH_DTYPE
    rng = np.random.default_rng(seed)
    sf = (1.2 ** np.arange(8)).astype(np.float32)
    K = np.array([f, f, w / 2 - 0.5, h / 2 - 0.5], np.float32)
    Rm = _rodrigues(np.array([0.01, -0.04, 0.02]))
    t = np.array([-baseline, 0.02, 0.03])
    RT = np.eye(4, dtype=np.float32); RT[:3, :3] = Rm; RT[:3, 3] = t
    far = rng.random(n) < far_frac
    Z = np.where(far, rng.uniform(60, 400, n), rng.uniform(1.0, 8.0, n))
    u = rng.uniform(30, w - 30, n); v = rng.uniform(30, h - 30, n)
    X = np.c_[(u - K[2]) / f * Z, (v - K[3]) / f * Z, Z]
    X2 = X @ Rm.T + t
    oct1 = rng.integers(0, 8, n); oct2 = np.clip(oct1 + rng.integers(-1, 2, n), 0, 7)
    k1 = np.zeros(n, KP_DTYPE); k2 = np.zeros(n, KP_DTYPE)
    k1["x"] = u + rng.normal(0, px_sigma, n) * sf[oct1]; k1["y"] = v + rng.normal(0, px_sigma, n) * sf[oct1]; k1["octave"] = oct1
    with np.errstate(all="ignore"):
        k2["x"] = f * X2[:, 0] / X2[:, 2] + K[2] + rng.normal(0, px_sigma, n) * sf[oct2]
        k2["y"] = f * X2[:, 1] / X2[:, 2] + K[3] + rng.normal(0, px_sigma, n) * sf[oct2]
    k2["octave"] = oct2
    wrong = rng.random(n) < outlier_frac
    k2["x"][wrong] = rng.uniform(0, w, int(wrong.sum())); k2["y"][wrong] = rng.uniform(0, h, int(wrong.sum()))
    perm = rng.permutation(n)                       # query keypoints are stored in another order than train keypoints
    k2s = np.zeros(n, KP_DTYPE); k2s[perm] = k2
    m = np.zeros(n, MATCH_DTYPE)
    m["trainIdx"] = np.arange(n); m["queryIdx"] = perm; m["distance"] = rng.integers(0, 60, n)
    m = m[rng.permutation(n)]
    gt = X.astype(np.float32); gt[wrong] = np.nan
    return dict(kps_train=k1, kps_query=k2s, matches=np.ascontiguousarray(m), K_train=K, K_query=K.copy(), RT=RT, sf_train=sf,
                sf_query=sf.copy(), xyz_gt=gt)


def synth_reloc_matches(seed, n=400, outlier_frac=0.4, px_sigma=0.8, w=640, h=480, f=525.0, backface_frac=0.1):
    """2D-3D matches as the relocaliser hands them to PnPSolver::solvePnPRansac: map points in front of a
    camera with a known pose, their pixels with noise, a share of wrong matches, and map-point normals (most facing the camera,
    some seen from behind so that the viewing-angle test rejects them).  Returns dict(p3d, p2d, normals, cam, pose_gt (4,4))."""
    rng = np.random.default_rng(seed)
    R = _rodrigues(rng.uniform(-0.4, 0.4, 3)); t = rng.uniform(-0.5, 0.5, 3) + np.array([0, 0, 0.5])
    cam = np.array([f, f, w / 2 - 0.5, h / 2 - 0.5], np.float32)
    Z = rng.uniform(2.0, 9.0, n)
    u = rng.uniform(20, w - 20, n); v = rng.uniform(20, h - 20, n)
    Xc = np.c_[(u - cam[2]) / f * Z, (v - cam[3]) / f * Z, Z]
    Xw = (Xc - t) @ R                      # Xc = R Xw + t
    p3d = Xw.astype(np.float32)
    Xc = p3d.astype(np.float64) @ R.T + t
    p2d = np.c_[f * Xc[:, 0] / Xc[:, 2] + cam[2], f * Xc[:, 1] / Xc[:, 2] + cam[3]] + rng.normal(0, px_sigma, (n, 2))
    wrong = rng.random(n) < outlier_frac
    p2d[wrong] = np.c_[rng.uniform(0, w, int(wrong.sum())), rng.uniform(0, h, int(wrong.sum()))]
    centre = -R.T @ t
    to_cam = centre - Xw
    to_cam /= np.linalg.norm(to_cam, axis=1)[:, None]
    normals = to_cam + rng.normal(0, 0.35, (n, 3))
    normals /= np.linalg.norm(normals, axis=1)[:, None]
    back = rng.random(n) < backface_frac
    normals[back] *= -1
    pose = np.eye(4); pose[:3, :3] = R; pose[:3, 3] = t
    return dict(p3d=p3d, p2d=p2d.astype(np.float32), normals=normals.astype(np.float32), cam=cam, pose_gt=pose)


def synth_new_points_scene(seed, n_kp=2000, n_nb=6, assigned_frac=0.4, w=640, h=480, f=525.0, flips=10, px_sigma=0.5, far_frac=0.1):
    """A new keyframe and n_nb covisible neighbour keyframes as MapManager::createNewPoints sees them: n_kp 3-D
    points in front of the keyframe (some too far for the parallax gate), every frame observes a random 70 % of them plus clutter, with
    octave-scaled pixel noise, descriptors = the point's descriptor with a few flipped bits, orientation turned with the camera roll;
    a share of the keypoints already has a map point (MODE_UNASSIGNED leaves them out: t_map / q_map).  Poses are global -> camera
    (pose_f2g); rt[f] = nb_f.pose_f2g * kf.pose_f2g.inv(); f12[f] = the fundamental matrix the matcher's epipolar gate uses
    (x_t^T F x_q = 0 for a train keypoint x_t of the keyframe and its match x_q in neighbour f), f32."""
    from . import KP_DTYPE
    rng = np.random.default_rng(seed)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    K = np.array([f, f, w / 2 - 0.5, h / 2 - 0.5], np.float32)
    Km = np.array([[f, 0, K[2]], [0, f, K[3]], [0, 0, 1]], np.float64)
    far = rng.random(n_kp) < far_frac
    Z = np.where(far, rng.uniform(80, 400, n_kp), rng.uniform(1.5, 8.0, n_kp))
    u = rng.uniform(30, w - 30, n_kp); v = rng.uniform(30, h - 30, n_kp)
    Xc = np.c_[(u - K[2]) / f * Z, (v - K[3]) / f * Z, Z]                  # in the keyframe's camera
    kf_pose = np.eye(4); kf_pose[:3, :3] = _rodrigues(rng.normal(0, 0.3, 3)); kf_pose[:3, 3] = rng.normal(0, 1.0, 3)   # global -> kf camera
    g2f = np.linalg.inv(kf_pose)
    Xw = Xc @ g2f[:3, :3].T + g2f[:3, 3]
    base_desc = rng.integers(0, 256, (n_kp, 32), dtype=np.uint8)
    base_angle = rng.uniform(0, 360, n_kp)
    base_oct = rng.integers(0, 7, n_kp)

    def observe(pose, roll_deg, frac):
        Xf = Xw @ pose[:3, :3].T + pose[:3, 3]
        with np.errstate(all="ignore"):
            px = f * Xf[:, 0] / Xf[:, 2] + K[2]; py = f * Xf[:, 1] / Xf[:, 2] + K[3]
        vis = (Xf[:, 2] > 0.3) & (px > 5) & (px < w - 5) & (py > 5) & (py < h - 5) & (rng.random(n_kp) < frac)
        ids = np.nonzero(vis)[0]
        n_cl = len(ids) // 4                                                # clutter: keypoints of nothing in common
        n = len(ids) + n_cl
        kp = np.zeros(n, KP_DTYPE)
        octv = np.clip(base_oct[ids] + rng.integers(-1, 2, len(ids)) * (rng.random(len(ids)) < 0.3), 0, 7)
        kp["x"][:len(ids)] = px[ids] + rng.normal(0, px_sigma, len(ids)) * sf[octv]
        kp["y"][:len(ids)] = py[ids] + rng.normal(0, px_sigma, len(ids)) * sf[octv]
        kp["octave"][:len(ids)] = octv
        kp["angle"][:len(ids)] = np.mod(base_angle[ids] + roll_deg + rng.normal(0, 4, len(ids)), 360)
        kp["x"][len(ids):] = rng.uniform(5, w - 5, n_cl); kp["y"][len(ids):] = rng.uniform(5, h - 5, n_cl)
        kp["octave"][len(ids):] = rng.integers(0, 8, n_cl); kp["angle"][len(ids):] = rng.uniform(0, 360, n_cl)
        kp["size"] = 31 * sf[kp["octave"]]; kp["class_id"] = -1
        desc = np.empty((n, 32), np.uint8)
        desc[:len(ids)] = base_desc[ids]
        fl = rng.integers(0, 256, (len(ids), flips))
        for j in range(flips):
            desc[np.arange(len(ids)), fl[:, j] >> 3] ^= (1 << (fl[:, j] & 7)).astype(np.uint8)
        desc[len(ids):] = rng.integers(0, 256, (n_cl, 32), dtype=np.uint8)
        perm = rng.permutation(n)
        point = np.concatenate([ids, np.full(n_cl, -1)])[perm]
        unassigned = np.sort(np.nonzero(rng.random(n) >= assigned_frac)[0]).astype(np.int32)
        return kp[perm], desc[perm], point, unassigned

    t_kps, t_desc, t_point, t_map = observe(kf_pose, 0.0, 0.9)
    q_kps, q_desc, q_point, q_map, rts, f12s, poses = [], [], [], [], [], [], []
    for i in range(n_nb):
        dpose = np.eye(4)
        dpose[:3, :3] = _rodrigues(rng.normal(0, 0.05, 3))
        dpose[:3, 3] = rng.normal(0, 0.25, 3) + np.array([0.3 * (1 if i % 2 else -1), 0, 0])
        pose = dpose @ kf_pose                                              # T_f = dpose: kf camera -> neighbour camera
        kp, de, pt, un = observe(pose, float(np.degrees(np.arctan2(dpose[1, 0], dpose[0, 0]))), 0.7)
        q_kps.append(kp); q_desc.append(de); q_point.append(pt); q_map.append(un)
        Rm, t = dpose[:3, :3], dpose[:3, 3]
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        Fm = (np.linalg.inv(Km).T @ tx @ Rm @ np.linalg.inv(Km)).T          # x_t^T F x_q = 0: epipolarLineSqDist(train, query, F12), misc.h:72-81
        f12s.append((Fm / np.abs(Fm).max()).astype(np.float32)); rts.append(dpose.astype(np.float32)); poses.append(pose)
    return dict(t_kps=t_kps, t_desc=t_desc, t_map=t_map, t_point=t_point, q_kps=q_kps, q_desc=q_desc, q_map=q_map, q_point=q_point,
                f12=np.array(f12s), rt=np.array(rts), K_kf=K, K_nb=np.tile(K, (n_nb, 1)), g2f_kf=g2f.astype(np.float32), sf_kf=sf, sf_nb=sf.copy(),
                min_desc_dist=100.0, ratio=0.6, scale_ratio_factor=float(np.float32(1.5) * np.float32(1.2)), points_gt=Xw, kf_pose=kf_pose, nb_poses=poses)
