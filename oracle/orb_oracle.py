"""TEST INFRASTRUCTURE ONLY — CPU oracle of the reference's ORB extractor.

Line-by-line restatement of /root/reference/src/featureextractors/ORBextractor.cpp (cited per function) on top of the
REAL third-party primitives the reference calls, so that only the reference's own control flow is restated:
  * OpenCV (un-vendored dependency of the reference; pinned here to the container's cv2 4.13.0 with IPP switched OFF, which
    is the open-source code path a distribution OpenCV build runs): cv2.GaussianBlur, cv2.resize(INTER_CUBIC),
    cv2.copyMakeBorder, cv2.FastFeatureDetector (FAST-9/16 + 3x3 NMS), cv2.fastAtan2
  * libstdc++ std::nth_element / std::partition for KeyPointsFilter::retainBest and libm cosf/sinf, through
    oracle/stl_helper.cpp
Parity pin: the reference ships no golden vectors for this path and cannot be compiled here (needs OpenCV C++
headers), so parity is pinned to these real library calls; tests/golden/orb_*.npz holds outputs of THIS oracle for
cross-machine regression.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import ctypes, os
import numpy as np
import cv2

cv2.ipp.setUseIPP(False)
cv2.setNumThreads(1)

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32
PATCH_SIZE, HALF_PATCH_SIZE, EDGE_THRESHOLD = 31, 15, 19          # ORBextractor.cpp:73-75
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])       # cv::KeyPoint, 28 bytes
PATTERN = np.loadtxt(os.path.join(HERE, "orb_bit_pattern_31.txt"), dtype=np.int32).reshape(512, 2)  # :156-414

_stl = None


def stl():
    global _stl
    if _stl is None:
        p = os.path.join(HERE, "_build", "liboracle_stl.so")
        if not os.path.exists(p):
            import oracle_py
            oracle_py.build_oracle()
        _stl = ctypes.CDLL(p)
        _stl.stl_retain_best.restype = ctypes.c_int
    return _stl


def cv_round(x):
    """cvRound(float/double): round half to even (SSE cvtss2si / lrint)."""
    return int(np.rint(x))


def retain_best(kps, n_points):
    """cv::KeyPointsFilter::retainBest, executed by the real libstdc++ (stl_helper.cpp)."""
    kps = np.array(kps, dtype=KP_DTYPE, copy=True)
    n = stl().stl_retain_best(kps.ctypes.data_as(ctypes.c_void_p), len(kps), int(n_points))
    return kps[:n].copy()


def umax_table():
    """ORBextractor.cpp:436-451."""
    umax = [0] * (HALF_PATCH_SIZE + 1)
    vmax = int(np.floor(f32(HALF_PATCH_SIZE) * np.sqrt(f32(2.0)) / f32(2) + f32(1)))
    vmin = int(np.ceil(f32(HALF_PATCH_SIZE) * np.sqrt(f32(2.0)) / f32(2)))
    hp2 = float(HALF_PATCH_SIZE * HALF_PATCH_SIZE)
    for v in range(vmax + 1):
        umax[v] = cv_round(np.sqrt(hp2 - v * v))
    v0 = 0
    for v in range(HALF_PATCH_SIZE, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax


UMAX = umax_table()


class Params:
    """precalculateParams, ORBextractor.cpp:466-514 (float arithmetic kept in f32 where the reference has it)."""

    def __init__(self, max_features=2000, n_levels=8, scale_factor=1.2, ini_th=20, min_th=7):
        self.max_features, self.n_levels = int(max_features), int(n_levels)
        self.scale_factor = f32(scale_factor)
        self.ini_th, self.min_th = ini_th, min_th
        sf = [f32(1.0)]
        for i in range(1, n_levels):
            sf.append(f32(sf[-1] * self.scale_factor))
        self.scale = sf
        self.inv_scale = [f32(f32(1.0) / s) for s in sf]
        factor = f32(f32(1.0) / self.scale_factor)
        n_desired = f32(f32(f32(self.max_features) * f32(f32(1) - factor)) /
                        f32(f32(1) - f32(np.power(np.float64(factor), np.float64(n_levels)))))
        self.n_per_level, s = [], 0
        for _ in range(n_levels - 1):
            self.n_per_level.append(cv_round(n_desired))
            s += self.n_per_level[-1]
            n_desired = f32(n_desired * factor)
        self.n_per_level.append(max(self.max_features - s, 0))

    def level_size(self, cols, rows, level):
        """ComputePyramid, ORBextractor.cpp:1369-1370."""
        sc = self.inv_scale[level]
        return cv_round(f32(f32(cols) * sc)), cv_round(f32(f32(rows) * sc))


def compute_pyramid(image, P, blur_first=True):
    """compute():1261-1266 + ComputePyramid:1355-1392. Returns the bordered buffers (level image = buf[19:-19,19:-19])."""
    if blur_first:
        img = cv2.GaussianBlur(image, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    else:
        img = image
    E = EDGE_THRESHOLD
    bufs = []
    for level in range(P.n_levels):
        w, h = P.level_size(image.shape[1], image.shape[0], level)
        if level != 0:
            prev = bufs[level - 1][E:-E, E:-E]
            lv = cv2.resize(prev, (w, h), interpolation=cv2.INTER_CUBIC)
        else:
            lv = img
        bufs.append(cv2.copyMakeBorder(lv, E, E, E, E, cv2.BORDER_REFLECT_101))
    return bufs


def cell_grid(P, level, cols0, rows0, cols, rows):
    """Grid geometry of ComputeKeyPoints_thread, ORBextractor.cpp:900-924."""
    image_ratio = f32(f32(cols0) / f32(rows0))
    n_desired = P.n_per_level[level]
    level_cols = int(np.sqrt(f32(f32(n_desired) / f32(f32(5) * image_ratio))))
    level_rows = int(f32(image_ratio * f32(level_cols)))
    min_bx = min_by = EDGE_THRESHOLD
    max_bx, max_by = cols - EDGE_THRESHOLD, rows - EDGE_THRESHOLD
    W, H = max_bx - min_bx, max_by - min_by
    if level_cols <= 0 or level_rows <= 0:
        raise ValueError("degenerate grid")
    cell_w = int(np.ceil(f32(f32(W) / f32(level_cols))))
    cell_h = int(np.ceil(f32(f32(H) / f32(level_rows))))
    n_cells = level_rows * level_cols
    nf_cell = int(np.ceil(f32(f32(n_desired) / f32(n_cells))))
    return dict(level_cols=level_cols, level_rows=level_rows, cell_w=cell_w, cell_h=cell_h, n_cells=n_cells,
                nf_cell=nf_cell, min_bx=min_bx, min_by=min_by, max_bx=max_bx, max_by=max_by, n_desired=n_desired)


def _fast(roi, th):
    det = cv2.FastFeatureDetector_create(int(th), True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = det.detect(np.ascontiguousarray(roi), None)
    out = np.zeros(len(kps), KP_DTYPE)
    for i, k in enumerate(kps):
        out[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
    return out


def ic_angle(img, step_img, x, y):
    """IC_Angle, ORBextractor.cpp:79-106. img: bordered buffer, (x, y) integer position inside it."""
    m01 = m10 = 0
    row = img[y].astype(np.int64)
    u = np.arange(-HALF_PATCH_SIZE, HALF_PATCH_SIZE + 1)
    m10 += int((u * row[x - HALF_PATCH_SIZE:x + HALF_PATCH_SIZE + 1]).sum())
    for v in range(1, HALF_PATCH_SIZE + 1):
        d = UMAX[v]
        uu = np.arange(-d, d + 1)
        plus = img[y + v, x - d:x + d + 1].astype(np.int64)
        minus = img[y - v, x - d:x + d + 1].astype(np.int64)
        m01 += v * int((plus - minus).sum())
        m10 += int((uu * (plus + minus)).sum())
    return f32(cv2.fastAtan2(float(f32(m01)), float(f32(m10))))


_DISC = None


def ic_angles(img, xs, ys):
    """IC_Angle for many keypoints at once: the same integer moments (exact, so the order of the sums does not matter) gathered
    with one fancy index per level instead of a Python loop per keypoint; fastAtan2 stays OpenCV's scalar function."""
    global _DISC
    if _DISC is None:
        dv, du = [], []
        for v in range(-HALF_PATCH_SIZE, HALF_PATCH_SIZE + 1):
            d = UMAX[abs(v)] if v != 0 else HALF_PATCH_SIZE
            for u in range(-d, d + 1):
                dv.append(v)
                du.append(u)
        _DISC = (np.array(dv, np.int64), np.array(du, np.int64))
    dv, du = _DISC
    xs, ys = np.asarray(xs, np.int64), np.asarray(ys, np.int64)
    vals = img[ys[:, None] + dv[None, :], xs[:, None] + du[None, :]].astype(np.int64)
    m01, m10 = (vals * dv).sum(1), (vals * du).sum(1)
    return np.array([cv2.fastAtan2(float(f32(a)), float(f32(b))) for a, b in zip(m01, m10)], f32)


FACTOR_PI = f32(np.pi / np.float64(f32(180.0)))   # (float)(CV_PI/180.f), ORBextractor.cpp:112


def orb_descriptors(buf, xs, ys, angles_deg):
    """computeOrbDescriptor, ORBextractor.cpp:113-153, vectorised over keypoints; f32 mul/add, no fused operations."""
    n = len(xs)
    desc = np.zeros((n, 32), np.uint8)
    if n == 0:
        return desc
    ang = (np.asarray(angles_deg, f32) * FACTOR_PI).astype(f32)
    a = np.empty(n, f32)
    b = np.empty(n, f32)
    stl().stl_sincosf_array(ang.ctypes.data_as(ctypes.c_void_p), n, a.ctypes.data_as(ctypes.c_void_p),
                            b.ctypes.data_as(ctypes.c_void_p))          # a = cosf(angle), b = sinf(angle)
    px = PATTERN[:, 0].astype(f32)[None, :]
    py = PATTERN[:, 1].astype(f32)[None, :]
    a_, b_ = a[:, None], b[:, None]
    rr = np.rint((px * b_).astype(f32) + (py * a_).astype(f32)).astype(np.int64)      # cvRound(x*b + y*a)
    cc = np.rint((px * a_).astype(f32) - (py * b_).astype(f32)).astype(np.int64)      # cvRound(x*a - y*b)
    vals = buf[np.asarray(ys)[:, None] + rr, np.asarray(xs)[:, None] + cc]            # (n, 512)
    bits = (vals[:, 0::2] < vals[:, 1::2]).astype(np.uint8)                            # (n, 256)
    desc = np.packbits(bits.reshape(n, 32, 8), axis=2, bitorder="little").reshape(n, 32)
    return desc


def compute_level_keypoints(buf, P, level, cols0, rows0, want_candidates=False):
    """ComputeKeyPoints_thread for one level, ORBextractor.cpp:899-1076. buf = bordered level buffer."""
    E = EDGE_THRESHOLD
    img = buf[E:-E, E:-E]
    rows, cols = img.shape
    g = cell_grid(P, level, cols0, rows0, cols, rows)
    LR, LC, cell_w, cell_h = g["level_rows"], g["level_cols"], g["cell_w"], g["cell_h"]
    n_cells, nf_cell, n_desired = g["n_cells"], g["nf_cell"], g["n_desired"]
    cell_kps = [[np.zeros(0, KP_DTYPE) for _ in range(LC)] for _ in range(LR)]
    cand = {}
    n_retain = np.zeros((LR, LC), np.int64)
    n_total = np.zeros((LR, LC), np.int64)
    no_more = np.zeros((LR, LC), bool)
    ini_x_col, ini_y_row = [0] * LC, [0] * LR
    n_no_more = n_distribute = 0
    hY = cell_h + 6
    for i in range(LR):
        iniY = g["min_by"] + i * cell_h - 3
        ini_y_row[i] = iniY
        if i == LR - 1:
            hY = g["max_by"] + 3 - iniY
            if hY <= 0:
                continue
        hX = cell_w + 6
        for j in range(LC):
            if i == 0:
                iniX = g["min_bx"] + j * cell_w - 3
                ini_x_col[j] = iniX
            else:
                iniX = ini_x_col[j]
            if j == LC - 1:
                hX = g["max_bx"] + 3 - iniX
                if hX <= 0:
                    continue
            if iniY + hY > rows or iniX + hX > cols:
                raise ValueError("cell outside the level image (cv::Mat::rowRange would assert)")
            roi = buf[E + iniY:E + iniY + hY, E + iniX:E + iniX + hX]
            k = _fast(roi, P.ini_th)
            if want_candidates:
                cand[(i, j)] = _fast(roi, P.min_th)
            if len(k) <= 3:
                k = cand[(i, j)] if want_candidates else _fast(roi, P.min_th)
            cell_kps[i][j] = k
            n_keys = len(k)
            n_total[i, j] = n_keys
            if n_keys > nf_cell:
                n_retain[i, j] = nf_cell
                no_more[i, j] = False
            else:
                n_retain[i, j] = n_keys
                n_distribute += nf_cell - n_keys
                no_more[i, j] = True
                n_no_more += 1
    while n_distribute > 0 and n_no_more < n_cells:                                   # :1013-1039
        n_new = int(f32(nf_cell) + np.ceil(f32(f32(n_distribute) / f32(n_cells - n_no_more))))
        n_distribute = 0
        for i in range(LR):
            for j in range(LC):
                if not no_more[i, j]:
                    if n_total[i, j] > n_new:
                        n_retain[i, j] = n_new
                        no_more[i, j] = False
                    else:
                        n_retain[i, j] = n_total[i, j]
                        n_distribute += n_new - n_total[i, j]
                        no_more[i, j] = True
                        n_no_more += 1
    scaled_patch = int(f32(PATCH_SIZE) * P.scale[level])
    out = []
    for i in range(LR):
        for j in range(LC):
            kc = retain_best(cell_kps[i][j], n_retain[i, j])
            if len(kc) > n_retain[i, j]:
                kc = kc[:n_retain[i, j]]
            kc = kc.copy()
            kc["x"] += f32(ini_x_col[j])
            kc["y"] += f32(ini_y_row[i])
            kc["octave"] = level
            kc["size"] = f32(scaled_patch)
            out.append(kc)
    kps = np.concatenate(out) if out else np.zeros(0, KP_DTYPE)
    if len(kps) > n_desired:
        kps = retain_best(kps, n_desired)[:n_desired].copy()
    if len(kps):                                                                       # computeOrientation :516
        kps["angle"] = ic_angles(buf, [E + cv_round(v) for v in kps["x"]], [E + cv_round(v) for v in kps["y"]])
    return (kps, cand, g, ini_x_col, ini_y_row) if want_candidates else kps


def extract(image, max_features=2000, n_levels=8, scale_factor=1.2, ini_th=20, min_th=7, blur_first=True,
            want_intermediates=False):
    """ORBextractor::compute (nthreads==1 path, :1271-1303). Returns (keypoints[KP_DTYPE], desc[N,32])."""
    assert image.dtype == np.uint8 and image.ndim == 2
    P = Params(max_features, n_levels, scale_factor, ini_th, min_th)
    bufs = compute_pyramid(image, P, blur_first)
    E = EDGE_THRESHOLD
    all_k, all_d, inter = [], [], dict(pyramid=bufs, levels=[])
    for level in range(n_levels):
        buf = bufs[level]
        if want_intermediates:
            kps, cand, g, ixc, iyr = compute_level_keypoints(buf, P, level, image.shape[1], image.shape[0], True)
            inter["levels"].append(dict(candidates=cand, grid=g, ini_x_col=ixc, ini_y_row=iyr, selected=kps.copy()))
        else:
            kps = compute_level_keypoints(buf, P, level, image.shape[1], image.shape[0])
        rows, cols = buf.shape[0] - 2 * E, buf.shape[1] - 2 * E
        keep = ~((kps["x"] < 19) | (kps["y"] < 19) | (kps["x"] > cols - 19) | (kps["y"] > rows - 19))   # :1124-1130
        kps = kps[keep]
        xs = np.rint(kps["x"]).astype(np.int64) + E
        ys = np.rint(kps["y"]).astype(np.int64) + E
        d = orb_descriptors(buf, xs, ys, kps["angle"])
        if level != 0:                                                                  # :1229
            sc = P.scale[level]
            kps["x"] = ((kps["x"] + f32(0.5)).astype(f32) * sc).astype(f32)
            kps["y"] = ((kps["y"] + f32(0.5)).astype(f32) * sc).astype(f32)
        all_k.append(kps)
        all_d.append(d)
    K = np.concatenate(all_k)
    D = np.concatenate(all_d) if len(K) else np.zeros((0, 32), np.uint8)
    return (K, D, inter) if want_intermediates else (K, D)


def synth_texture(seed=1234, size=2048):
    """sum of 5 octaves of seeded block noise, normalised to [0,255] (FAST-dense corners)."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((size, size), np.float64)
    amp = 1.0
    for blk in (64, 32, 16, 8, 4):
        n = size // blk
        acc += amp * np.kron(rng.random((n, n)), np.ones((blk, blk)))
        amp *= 0.5
    acc -= acc.min()
    return (acc / acc.max() * 255).astype(np.uint8)


_TEX = {}


def synth_frame(idx, w=640, h=480, seed=1234):
    """Frame idx of the synthetic clip: a perspective view of the texture along a smooth camera path."""
    if seed not in _TEX:
        _TEX[seed] = synth_texture(seed)
    tex = _TEX[seed]
    t = idx * 0.02
    c, s = np.cos(0.15 * np.sin(t)), np.sin(0.15 * np.sin(t))
    zoom = 1.6 + 0.2 * np.sin(0.7 * t)
    Hm = np.array([[c * zoom, -s * zoom, 300 + 120 * t], [s * zoom, c * zoom, 400 + 40 * np.sin(t)],
                   [1e-4 * np.sin(t), 5e-5, 1.0]])
    return cv2.warpPerspective(tex, Hm, (w, h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                               borderMode=cv2.BORDER_REFLECT_101)
