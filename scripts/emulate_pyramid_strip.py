"""CPU emulation of the index arithmetic of blur7_strip_kernel / resize_cubic_strip_kernel (ucoslam-cv3_b200/csrc/orb.cu),
statement by statement (thread loops as Python loops, __byte_perm / __funnelshift_r as integer functions), checked against
cv2.GaussianBlur / cv2.resize / cv2.copyMakeBorder on small images.  Design check only: the kernels themselves are compared with
the oracle by tests/test_orb_gpu.py on the GPU.   python scripts/emulate_pyramid_strip.py"""
import sys, os
import numpy as np
import cv2

cv2.ipp.setUseIPP(False)   # as oracle/orb_oracle.py: the open-source OpenCV code path is the contract (IPP's cubic resize rounds differently)
E = 19
BS_R, BS_STEPS, BS_MAXW = 29, 35, 256
M32 = 0xffffffff


def reflect101(p, n):
    if 0 <= p < n:
        return p
    if n == 1:
        return 0
    while not (0 <= p < n):
        p = -p if p < 0 else 2 * (n - 1) - p
    return p


def byte_perm(a, b, s):
    by = [(a >> (8 * i)) & 0xff for i in range(4)] + [(b >> (8 * i)) & 0xff for i in range(4)]
    r = 0
    for i in range(4):
        n = (s >> (4 * i)) & 0xf
        assert n < 8
        r |= by[n] << (8 * i)
    return r


def funnelshift_r(lo, hi, sh):
    return (((hi << 32) | lo) >> sh) & M32


def store_word_rows(buf, col, h, y, word):
    rows = [y + E]
    if 1 <= y <= E:
        rows.append(E - y)
    if h - 1 - E <= y <= h - 2:
        rows.append(2 * (h - 1) - y + E)
    for r in rows:
        for j in range(4):
            buf[r, col + j] = (word >> (8 * j)) & 0xff


def blur_strip(img, do_blur=True, aligned=True):
    h, w = img.shape
    pitch = (w + 2 * E + 63) & ~63
    out = np.full((h + 2 * E, pitch), 0xAA, np.uint8)
    nwords = (w + 2 * E + 3) >> 2
    ntx = (nwords + BS_MAXW - 1) // BS_MAXW
    wpt = (nwords + ntx - 1) // ntx
    spw = wpt + 2
    for bx in range(ntx):
        for by in range((h + BS_R - 1) // BS_R):
            k0 = bx * wpt
            nk = min(wpt, nwords - k0)
            y0 = by * BS_R
            tile = np.zeros((BS_STEPS, spw * 4), np.uint8)
            vb0 = 4 * (k0 - 6)
            s_lo = s_hi = nk + 2
            if aligned:
                s_lo = min(max(0, 6 - k0), nk + 2)
                s_hi = max(s_lo, min(nk + 2, ((w - 4 - vb0) >> 2) + 1))
            n_left = 4 * s_lo
            n_edge = n_left + 4 * (nk + 2 - s_hi)
            for r in range(BS_STEPS):
                row = img[reflect101(y0 - 3 + r, h)]
                for s in range(s_lo, s_hi):
                    assert 0 <= vb0 + 4 * s and vb0 + 4 * s + 3 <= w - 1
                    tile[r, 4 * s:4 * s + 4] = row[vb0 + 4 * s:vb0 + 4 * s + 4]
                for e in range(n_edge):
                    bi = e if e < n_left else 4 * s_hi + (e - n_left)
                    tile[r, bi] = row[reflect101(vb0 + bi, w)]
            tw = tile.view(np.uint32)
            for tid in range(nk):
                col = 4 * (k0 + tid)
                if not do_blur:
                    for r in range(BS_R):
                        if y0 + r >= h:
                            break
                        store_word_rows(out, col, h, y0 + r, funnelshift_r(int(tw[r + 3, tid + 1]), int(tw[r + 3, tid + 2]), 8))
                    continue
                hw = [[0] * 4 for _ in range(7)]
                for g in range(BS_STEPS // 7):
                    for u in range(7):
                        step = g * 7 + u
                        W0, W1, W2 = int(tw[step, tid]), int(tw[step, tid + 1]), int(tw[step, tid + 2])
                        p2 = byte_perm(W0, 0, 0x4342); p3 = byte_perm(W0, W1, 0x0403) & 0x00ff00ff; p4 = byte_perm(W1, 0, 0x4140)
                        p5 = byte_perm(W1, 0, 0x4241); p6 = byte_perm(W1, 0, 0x4342); p7 = byte_perm(W1, W2, 0x0403) & 0x00ff00ff
                        p8 = byte_perm(W2, 0, 0x4140); p9 = byte_perm(W2, 0, 0x4241); p10 = byte_perm(W2, 0, 0x4342)
                        A = (18 * (p2 + p8) + 34 * (p3 + p7) + 48 * (p4 + p6) + 56 * p5) & M32
                        B = (18 * (p4 + p10) + 34 * (p5 + p9) + 48 * (p6 + p8) + 56 * p7) & M32
                        hw[u] = [A & 0xffff, A >> 16, B & 0xffff, B >> 16]
                        yo = y0 + step - 6
                        if (g > 0 or u == 6) and yo < h:
                            v = [(18 * (hw[(u + 1) % 7][j] + hw[u][j]) + 34 * (hw[(u + 2) % 7][j] + hw[(u + 6) % 7][j]) +
                                  48 * (hw[(u + 3) % 7][j] + hw[(u + 5) % 7][j]) + 56 * hw[(u + 4) % 7][j] + 32768) & M32 for j in range(4)]
                            assert all(x < (1 << 24) for x in v)
                            word = byte_perm(byte_perm(v[0], v[1], 0x0062), byte_perm(v[2], v[3], 0x0062), 0x5410)
                            store_word_rows(out, col, h, yo, word)
    return out[:, :w + 2 * E]


def cubic_table(dst, src):
    inv_scale = dst / src
    scale = 1.0 / inv_scale
    ofs, coef = [], []
    for d in range(dst):
        fx = np.float32((d + 0.5) * scale - 0.5)
        sx = int(np.floor(fx))
        fx = np.float32(fx - np.float32(sx))
        A = np.float32(-0.75)
        one = np.float32(1)
        c0 = ((A * (fx + one) - np.float32(5) * A) * (fx + one) + np.float32(8) * A) * (fx + one) - np.float32(4) * A
        c1 = ((A + np.float32(2)) * fx - (A + np.float32(3))) * fx * fx + one
        c2 = ((A + np.float32(2)) * (one - fx) - (A + np.float32(3))) * (one - fx) * (one - fx) + one
        c3 = one - c0 - c1 - c2
        ofs.append(sx)
        coef.append([int(np.rint(np.float32(c) * np.float32(2048))) for c in (c0, c1, c2, c3)])
    return ofs, coef


def resize_strip(srcb, sw, sh, dw, dh):
    """srcb: bordered source level (sh + 38, >= sw + 38); returns the bordered destination level."""
    RS2_TW, RS2_TH = 64, 32
    ox, cx = cubic_table(dw, sw)
    oy, cy = cubic_table(dh, sh)
    pitch = (dw + 2 * E + 63) & ~63
    out = np.full((dh + 2 * E, pitch), 0xAA, np.uint8)
    wext = (dw + 2 * E + 3) & ~3
    vec_limit = (dw // 8) * 8
    src = srcb[E:, E:].astype(np.int64)
    f32 = np.float32
    for bx in range((dw + 2 * E + RS2_TW - 1) // RS2_TW):
        for by in range((dh + RS2_TH - 1) // RS2_TH):
            c0, y0 = bx * RS2_TW, by * RS2_TH
            ny = min(RS2_TH, dh - y0)
            r_lo = min(max(oy[y0] - 1, 0), sh - 1)
            r_hi = min(max(oy[y0 + ny - 1] + 2, 0), sh - 1)
            nr = r_hi - r_lo + 1
            assert nr <= 72
            sr = np.zeros((72, RS2_TW + 4), np.int64)
            fix = np.zeros(RS2_TW, np.uint8)
            cl = min(c0 + RS2_TW, wext) - 1
            pa, pb = reflect101(c0 - E, dw), reflect101(cl - E, dw)
            pmin, pmax = min(pa, pb), max(pa, pb)
            if c0 <= E <= cl:
                pmin = 0
            if c0 <= dw - 1 + E <= cl:
                pmax = dw - 1
            xlo, xhi = min(max(ox[pmin] - 1, 0), sw - 1), min(max(ox[pmax] + 2, 0), sw - 1)
            a0 = (xlo + E) & ~3
            nw = ((xhi + E - a0) >> 2) + 1
            assert nw <= 36
            s_src = np.full((72, 144), -1, np.int64)
            for r in range(nr):
                s_src[r, :4 * nw] = srcb[r_lo + E + r, a0:a0 + 4 * nw]
            for x in range(RS2_TW):
                c = c0 + x
                if c >= wext:
                    continue
                px = reflect101(c - E, dw)
                sx, a = ox[px], cx[px]
                q = [min(max(sx - 1 + k, 0), sw - 1) + E - a0 for k in range(4)]
                assert min(q) >= 0 and max(q) < 4 * nw
                fix[x] = px >= vec_limit
                for r in range(nr):
                    sr[r, x] = sum(int(s_src[r, q[k]]) * a[k] for k in range(4))
            for i in range(16 * RS2_TH):
                y, gx = i >> 4, i & 15
                if y >= ny or c0 + 4 * gx >= wext:
                    continue
                dy = y0 + y
                sy, b = oy[dy], cy[dy]
                R = [sr[min(max(sy - 1 + k, 0), sh - 1) - r_lo, 4 * gx:4 * gx + 4] for k in range(4)]
                scale = f32(1.0) / (f32(2048.0) * f32(2048.0))
                bf = [f32(f32(b[k]) * scale) for k in range(4)]
                word = 0
                for j in range(4):
                    if fix[4 * gx + j]:
                        acc = sum(int(R[k][j]) * b[k] for k in range(4))
                        v = (acc + (1 << 21)) >> 22
                    else:
                        t = f32(f32(R[3][j]) * bf[3])
                        t = f32(f32(f32(R[2][j]) * bf[2]) + t)
                        t = f32(f32(f32(R[1][j]) * bf[1]) + t)
                        t = f32(f32(f32(R[0][j]) * bf[0]) + t)
                        v = int(np.rint(t))
                    word |= min(max(v, 0), 255) << (8 * j)
                store_word_rows(out, c0 + 4 * gx, dh, dy, word)
    return out[:, :dw + 2 * E]


def main():
    rng = np.random.default_rng(7)
    ok = True
    for (w, h) in [(64, 64), (97, 70), (130, 66), (640 // 4 + 3, 90)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ref = cv2.copyMakeBorder(cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101), E, E, E, E, cv2.BORDER_REFLECT_101)
        for aligned in (True, False):
            got = blur_strip(img, True, aligned)
            n = int((got != ref).sum())
            print("blur %dx%d aligned=%d: %d px differ" % (w, h, aligned, n)); ok &= n == 0
        got = blur_strip(img, False)
        n = int((got != cv2.copyMakeBorder(img, E, E, E, E, cv2.BORDER_REFLECT_101)).sum())
        print("copy %dx%d: %d px differ" % (w, h, n)); ok &= n == 0
    for (sw, sh, sc) in [(120, 90, 1.2), (131, 77, 1.5), (160, 120, 2.0), (101, 99, 1.3)]:
        img = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        dw, dh = int(np.rint(np.float32(sw) * np.float32(1.0 / sc))), int(np.rint(np.float32(sh) * np.float32(1.0 / sc)))
        ref = cv2.copyMakeBorder(cv2.resize(img, (dw, dh), interpolation=cv2.INTER_CUBIC), E, E, E, E, cv2.BORDER_REFLECT_101)
        got = resize_strip(cv2.copyMakeBorder(img, E, E, E, E, cv2.BORDER_REFLECT_101), sw, sh, dw, dh)
        n = int((got != ref).sum())
        print("resize %dx%d -> %dx%d: %d px differ" % (sw, sh, dw, dh, n)); ok &= n == 0
    print("OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
