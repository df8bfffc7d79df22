"""TEST INFRASTRUCTURE ONLY -- oracle of the feature front-end ``extract_points`` (slam.py:112-122).

``extract_points_cv2`` makes the reference's own third-party calls (real opencv-python; the reference's
``cv2.KeyPoint(..., _size=20)`` kwarg is spelled ``size`` on cv2 >= 4.5.3).  The numpy functions below restate
what those calls compute, with the float-operation order that reproduces opencv-python 4.13 bit for bit; they
are the readable spec the CUDA kernels in ``openwork_b200/csrc/frontend.cu`` follow, and
``tests/test_frontend.py`` pins them against the real cv2 on the CPU."""
from __future__ import annotations

import numpy as np

f32 = np.float32


def extract_points_cv2(image, max_corners=3000, quality_level=0.01, min_distance=3):
    """slam.py:112-122 with the real cv2: returns (kps f64[N,2], des u8[N,32])."""
    import cv2
    orb = cv2.ORB_create()
    gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY) if image.ndim == 3 else image
    pts = cv2.goodFeaturesToTrack(gray, max_corners, qualityLevel=quality_level, minDistance=min_distance)
    if pts is None:
        return np.zeros((0, 2)), np.zeros((0, 32), np.uint8)
    kps = [cv2.KeyPoint(x=float(pt[0][0]), y=float(pt[0][1]), size=20) for pt in pts]
    kps, des = orb.compute(gray, kps)
    if des is None:
        return np.zeros((0, 2)), np.zeros((0, 32), np.uint8)
    return np.array([(kp.pt[0], kp.pt[1]) for kp in kps]), des


def _fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def _reflect_pad(a, r):
    return np.pad(a, r, mode="reflect")          # BORDER_REFLECT_101


def bgr2gray(img):
    b, g, r = [img[..., i].astype(np.int64) for i in range(3)]
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def min_eig_map(gray):
    """cv2.cornerMinEigenVal(gray, blockSize=3, ksize=3).  The 3x3 box sums are taken directly in double; OpenCV
    keeps running sums, which differs in the last bit of cov_xy for ~1e-5 of the pixels."""
    sc = 1.0 / (255.0 * 4 * 3)
    k0, k1 = f32(2 * sc), f32(sc)
    p = _reflect_pad(gray, 1).astype(f32)
    r = p[:, 2:] - p[:, :-2]
    dx = _fma(r[:-2] + r[2:], k1, (k0 * r[1:-1]).astype(f32))
    q = _fma(p[:, 2:], k1, _fma(p[:, 1:-1], k0, (k1 * p[:, :-2]).astype(f32)))
    dy = (q[2:] - q[:-2]).astype(f32)
    H, W = gray.shape

    def box(a):
        b = _reflect_pad(a, 1).astype(np.float64)
        return sum(b[i:i + H, j:j + W] for i in range(3) for j in range(3)).astype(f32)
    a = (box((dx * dx).astype(f32)) * f32(0.5)).astype(f32)
    b = box((dx * dy).astype(f32))
    c = (box((dy * dy).astype(f32)) * f32(0.5)).astype(f32)
    d = (a - c).astype(f32)
    return ((a + c).astype(f32) - np.sqrt(((d * d).astype(f32) + (b * b).astype(f32)).astype(f32)).astype(f32)).astype(f32)


def good_features_to_track(gray, max_corners=3000, quality_level=0.01, min_distance=3):
    """cv2.goodFeaturesToTrack restated: threshold at max*quality, 3x3 local maxima off the border, sort by
    (value, address) descending, greedy min-distance filter on a grid.  Returns f32[N,2] (x, y)."""
    e = min_eig_map(gray)
    H, W = e.shape
    thr = f32(float(e.max()) * quality_level)
    e = np.where(e > thr, e, f32(0))
    pad = np.pad(e, 1, mode="constant", constant_values=-np.inf)
    dil = np.max([pad[i:i + H, j:j + W] for i in range(3) for j in range(3)], axis=0)
    cand = (e != 0) & (e == dil)
    cand[0, :] = cand[-1, :] = False
    cand[:, 0] = cand[:, -1] = False
    ys, xs = np.nonzero(cand)
    order = np.lexsort((-(ys * W + xs), -e[ys, xs].astype(np.float64)))
    cell = int(round(min_distance))
    gw, gh = (W + cell - 1) // cell, (H + cell - 1) // cell
    grid, out, md2 = {}, [], float(min_distance) ** 2
    for i in order:
        x, y = int(xs[i]), int(ys[i])
        xc, yc = x // cell, y // cell
        good = True
        for cy in range(max(0, yc - 1), min(gh - 1, yc + 1) + 1):
            for cx in range(max(0, xc - 1), min(gw - 1, xc + 1) + 1):
                for px, py in grid.get((cy, cx), ()):
                    if (x - px) ** 2 + (y - py) ** 2 < md2:
                        good = False
        if good:
            grid.setdefault((yc, xc), []).append((x, y))
            out.append((x, y))
            if len(out) == max_corners:
                break
    return np.array(out, dtype=f32).reshape(-1, 2)


def gaussian7_sigma2():
    x = np.arange(7) - 3.0
    k = np.exp(-(x * x) / 8.0)
    return (k * (1.0 / k.sum())).astype(f32)


def blur_7x7_sigma2(gray):
    """What ORB.compute samples: fp32 separable 7x7 Gaussian (rows: sequential FMAs; columns: centre first,
    symmetric pairs), rounded half-to-even to u8."""
    k = gaussian7_sigma2()
    H, W = gray.shape
    p = _reflect_pad(gray, 3).astype(f32)
    rows = (k[0] * p[:, 0:W]).astype(f32)
    for j in range(1, 7):
        rows = _fma(p[:, j:j + W], k[j], rows)
    out = (k[3] * rows[3:3 + H]).astype(f32)
    for j in (1, 2, 3):
        out = _fma((rows[3 - j:3 - j + H] + rows[3 + j:3 + j + H]).astype(f32), k[3 - j], out)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def orb_compute(gray, pts, pattern):
    """cv2.ORB_create().compute(gray, [KeyPoint(x, y, size=20)]): 31-pixel border filter, then 256 tests on the
    blurred image at the (already rotated + rounded) pattern offsets.  Returns (kps f64[N,2], des u8[N,32])."""
    H, W = gray.shape
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    keep = (pts[:, 0] >= 31) & (pts[:, 0] < W - 31) & (pts[:, 1] >= 31) & (pts[:, 1] < H - 31)
    pts = pts[keep]
    blur = blur_7x7_sigma2(gray)
    cx, cy = np.rint(pts[:, 0]).astype(int), np.rint(pts[:, 1]).astype(int)
    ix, iy = pattern[:, 0].astype(int), pattern[:, 1].astype(int)
    v = blur[cy[:, None] + iy[None, :], cx[:, None] + ix[None, :]].astype(np.int32)
    bits = (v[:, 0::2] < v[:, 1::2]).astype(np.uint8)
    return pts, np.packbits(bits, axis=1, bitorder="little")


def synthetic_image(seed, height=480, width=640, channels=3, sigma=1.3):
    """Textured test image (smoothed noise + a few rectangles), BGR u8."""
    import cv2
    rng = np.random.default_rng(seed)
    img = (rng.random((height, width, 3)) * 255).astype(np.uint8)
    img = cv2.GaussianBlur(img, (0, 0), sigma)
    img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    for _ in range(12):
        x0, y0 = int(rng.integers(0, width - 60)), int(rng.integers(0, height - 60))
        cv2.rectangle(img, (x0, y0), (x0 + int(rng.integers(10, 60)), y0 + int(rng.integers(10, 60))),
                      [int(v) for v in rng.integers(0, 256, 3)], -1)
    return img if channels == 3 else img[..., 0].copy()
