"""Seeded synthetic inputs: textured grayscale frames, stereo pairs,
streams and bundle-adjustment problems.  Pure numpy; used by tests and bench.py."""
import numpy as np


def _box_blur3(img):
    p = np.pad(img, 1, mode="edge").astype(np.float32)
    acc = np.zeros_like(img, dtype=np.float32)
    for dy in range(3):
        for dx in range(3):
            acc += p[dy:dy + img.shape[0], dx:dx + img.shape[1]]
    return acc / 9.0


def frame(width, height, seed=0, num_rects=None, noise_sigma=3.0):
    """u8 gray frame: random-contrast rectangles + checker patches + noise, lightly
    smoothed, so that per-cell FAST at threshold 20 yields several times the target
    number of keypoints on every pyramid level."""
    rng = np.random.default_rng(seed)
    img = np.full((height, width), 110.0, np.float32)
    n = num_rects if num_rects is not None else max(200, (width * height) // 900)
    xs = rng.integers(0, width, n); ys = rng.integers(0, height, n)
    ws = rng.integers(4, 48, n); hs = rng.integers(4, 48, n)
    vals = rng.integers(20, 236, n)
    for x, y, w, h, v in zip(xs, ys, ws, hs, vals):
        img[y:y + h, x:x + w] = v
    m = max(40, n // 12)
    xs = rng.integers(0, max(1, width - 64), m); ys = rng.integers(0, max(1, height - 64), m)
    for x, y in zip(xs, ys):
        s = int(rng.integers(3, 9)); k = int(rng.integers(3, 8))
        a, b = rng.integers(20, 120), rng.integers(136, 236)
        yy, xx = np.mgrid[0:s * k, 0:s * k]
        patch = np.where(((yy // s) + (xx // s)) % 2 == 0, a, b).astype(np.float32)
        ph, pw = img[y:y + s * k, x:x + s * k].shape
        img[y:y + ph, x:x + pw] = patch[:ph, :pw]
    img = _box_blur3(img)
    img += rng.normal(0.0, noise_sigma, img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def shifted(img, dx, dy=0):
    """Stream helper: integer roll (exact ground truth for matching tests)."""
    return np.roll(np.roll(img, dy, axis=0), dx, axis=1)
