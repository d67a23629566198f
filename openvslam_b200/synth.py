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


# ------------------------------------------------------------------ bundle-adjustment problems
def _rot(axis_angle):
    a = np.asarray(axis_angle, np.float64)
    th = np.linalg.norm(a)
    if th < 1e-12:
        return np.eye(3)
    k = a / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def project(cam, pose, pw):
    """cam: dict(model, fx, fy, cx, cy, focal_x_baseline, cols, rows); pose: 12 (R row-major, t)."""
    R = np.asarray(pose[:9]).reshape(3, 3); t = np.asarray(pose[9:])
    pc = pw @ R.T + t
    if cam["model"] == "equirectangular":
        L = np.linalg.norm(pc, axis=-1)
        theta = np.arctan2(pc[..., 0], pc[..., 2]); phi = -np.arcsin(pc[..., 1] / L)
        return np.stack([cam["cols"] * (0.5 + theta / (2 * np.pi)), cam["rows"] * (0.5 - phi / np.pi)], -1), pc
    x = cam["fx"] * pc[..., 0] / pc[..., 2] + cam["cx"]
    y = cam["fy"] * pc[..., 1] / pc[..., 2] + cam["cy"]
    return np.stack([x, y, x - cam["focal_x_baseline"] / pc[..., 2]], -1), pc


def ba_problem(num_free=50, num_fixed=10, num_landmarks=20000, model="equirectangular", seed=0, stereo=False,
               pixel_sigma=1.0, outlier_frac=0.05, pose_noise=(0.01, 0.05), point_noise=0.05, obs_range=(2, 8)):
    """Local-BA problem in the layout local_bundle_adjuster::optimize builds its graph from
    (50 KF on a 10 m arc, 20k landmarks in a 20 m shell, ~5 observations per
    landmark, pixel noise sigma 1, 5% outliers).  Observations are grouped by landmark, as the
    reference adds them.  Returns a dict of numpy arrays (ground truth included)."""
    rng = np.random.default_rng(seed)
    K = num_free + num_fixed
    if model == "equirectangular":
        cam = dict(model=model, fx=0.0, fy=0.0, cx=0.0, cy=0.0, focal_x_baseline=0.0, cols=1920.0, rows=960.0)
    else:
        cam = dict(model="perspective", fx=718.856, fy=718.856, cx=607.19, cy=185.21, focal_x_baseline=386.1448, cols=1241.0, rows=376.0)
    poses = np.zeros((K, 12))
    centers = np.zeros((K, 3))
    for k in range(K):
        if model == "equirectangular":
            a = (k / max(K - 1, 1)) * 2.0  # 10 m arc of radius 5 m
            c = np.array([5 * np.cos(a), 0.1 * rng.standard_normal(), 5 * np.sin(a)])
            R = _rot([0.05 * rng.standard_normal(), a + 0.1 * rng.standard_normal(), 0.05 * rng.standard_normal()])
        else:
            c = np.array([0.05 * rng.standard_normal(), 0.02 * rng.standard_normal(), 0.4 * k])
            R = _rot(0.02 * rng.standard_normal(3))
        centers[k] = c
        poses[k, :9] = R.reshape(-1); poses[k, 9:] = -R @ c
    if model == "equirectangular":
        d = rng.standard_normal((num_landmarks, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        points = d * rng.uniform(8, 20, (num_landmarks, 1))
    else:
        z = rng.uniform(6, 60, num_landmarks) + 0.4 * K
        points = np.stack([rng.uniform(-0.7, 0.7, num_landmarks) * z * 0.5, rng.uniform(-0.2, 0.2, num_landmarks) * z * 0.5, z], 1)
    obs_kf, obs_lm, obs_xy, obs_xr, inv_s, is_outlier = [], [], [], [], [], []
    for l in range(num_landmarks):
        m = int(rng.integers(obs_range[0], obs_range[1] + 1))
        kfs = rng.choice(K, size=min(m, K), replace=False)
        for k in np.sort(kfs):
            uv, pc = project(cam, poses[k], points[l])
            if model != "equirectangular" and (pc[2] < 1.0 or not (0 <= uv[0] < cam["cols"] and 0 <= uv[1] < cam["rows"])):
                continue
            noise = rng.standard_normal(3) * pixel_sigma
            out = rng.random() < outlier_frac
            if out:
                noise[:2] += rng.choice([-1, 1], 2) * rng.uniform(15, 40, 2)
            level = int(rng.integers(0, 8))
            obs_kf.append(k); obs_lm.append(l)
            obs_xy.append((uv[0] + noise[0], uv[1] + noise[1]))
            obs_xr.append(uv[2] + noise[2] if (stereo and model != "equirectangular" and rng.random() < 0.8) else -1.0)
            inv_s.append(1.0 / (1.2 ** level) ** 2)
            is_outlier.append(out)
    # initial estimates: perturb the free poses and all landmarks
    poses0 = poses.copy()
    fixed = np.zeros(K, np.uint8); fixed[num_free:] = 1
    for k in range(num_free):
        R = poses[k, :9].reshape(3, 3); t = poses[k, 9:]
        dR = _rot(pose_noise[0] * rng.standard_normal(3))
        poses0[k, :9] = (dR @ R).reshape(-1); poses0[k, 9:] = dR @ t + pose_noise[1] * rng.standard_normal(3)
    points0 = points + point_noise * rng.standard_normal(points.shape)
    return dict(cam=cam, setup_is_mono=not stereo, poses_gt=poses, points_gt=points, poses=poses0, points=points0, fixed=fixed,
                obs_kf=np.array(obs_kf, np.int32), obs_lm=np.array(obs_lm, np.int32), obs_xy=np.array(obs_xy, np.float32),
                obs_xr=np.array(obs_xr, np.float32), inv_sigma_sq=np.array(inv_s, np.float32), is_outlier=np.array(is_outlier))


def pose_problem(num_points=2000, model="perspective", seed=0, stereo=True, pixel_sigma=1.0, outlier_frac=0.1, pose_noise=(0.02, 0.1)):
    """Motion-only problem for pose_optimizer::optimize: one frame, its matched landmarks."""
    p = ba_problem(num_free=1, num_fixed=0, num_landmarks=num_points, model=model, seed=seed, stereo=stereo,
                   pixel_sigma=pixel_sigma, outlier_frac=outlier_frac, pose_noise=pose_noise, point_noise=0.0, obs_range=(1, 1))
    p["pts_w"] = p["points_gt"][p["obs_lm"]]
    return p


def reprojection_chi2(cam, poses, points, obs_kf, obs_lm, obs_xy, obs_xr, inv_sigma_sq, mask=None):
    """Independent (numpy) evaluation of sum_i inv_sigma_sq_i * |obs_i - project_i|^2 over `mask`."""
    total = 0.0
    idx = np.arange(len(obs_kf)) if mask is None else np.flatnonzero(mask)
    for k in np.unique(obs_kf[idx]):
        sel = idx[obs_kf[idx] == k]
        uv, _ = project(cam, poses[k], points[obs_lm[sel]])
        e = obs_xy[sel].astype(np.float64) - uv[:, :2]
        c = (e ** 2).sum(1)
        if cam["model"] != "equirectangular" and obs_xr is not None:
            st = obs_xr[sel] >= 0
            c = c + np.where(st, (obs_xr[sel].astype(np.float64) - uv[:, 2]) ** 2, 0.0)
        total += float((c * inv_sigma_sq[sel]).sum())
    return total


def triangulation_problem(n=1500, seed=0, n_nodes=40, stereo_frac=0.1):
    """Two keyframes observing the same 3-D points (perspective bearings), for robust::match_for_triangulation:
    descriptors (a few bits flipped between the views, plus unrelated distractors), bearings, octaves, angles,
    landmark / stereo flags, BoW node ids (the same node for both views of a point, except for some corrupted ones),
    the essential matrix E_12 (b1' E_12 b2 = 0) and the bearing of camera centre 1 seen from keyframe 2."""
    rng = np.random.default_rng(seed)
    X1 = np.stack([rng.uniform(-6, 6, n), rng.uniform(-4, 4, n), rng.uniform(4, 20, n)], 1)       # in camera-1 coordinates
    ang = 0.05
    R21 = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t21 = np.array([-0.8, 0.05, 0.1])
    X2 = X1 @ R21.T + t21
    R12 = R21.T; t12 = -R21.T @ t21
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    E12 = tx @ R12
    nd = n // 3                                                    # distractors per keyframe
    b1 = np.concatenate([X1, np.stack([rng.uniform(-6, 6, nd), rng.uniform(-4, 4, nd), rng.uniform(4, 20, nd)], 1)])
    b2 = np.concatenate([X2, np.stack([rng.uniform(-6, 6, nd), rng.uniform(-4, 4, nd), rng.uniform(4, 20, nd)], 1)])
    b1 += rng.normal(0, 2e-3, b1.shape) * b1[:, 2:3]                # pixel-level noise: some pairs fall outside the 0.2 deg band
    b2 += rng.normal(0, 2e-3, b2.shape) * b2[:, 2:3]
    b1 /= np.linalg.norm(b1, axis=1, keepdims=True); b2 /= np.linalg.norm(b2, axis=1, keepdims=True)
    base = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    d1 = np.concatenate([base, rng.integers(0, 256, (nd, 32), dtype=np.uint8)])
    d2v = base.copy()
    for _ in range(3):                                             # up to 24 flipped bits, many pairs tie at equal distances
        byte = rng.integers(0, 32, n); d2v[np.arange(n), byte] ^= rng.integers(0, 256, n).astype(np.uint8)
    d2 = np.concatenate([d2v, rng.integers(0, 256, (nd, 32), dtype=np.uint8)])
    node = (base[:, 0].astype(np.int32) * 7 + base[:, 1]) % n_nodes
    node1 = np.concatenate([node, rng.integers(0, n_nodes, nd)]).astype(np.int32)
    node2 = np.concatenate([node, rng.integers(0, n_nodes, nd)]).astype(np.int32)
    bad = rng.random(n) < 0.08
    node2[:n][bad] = rng.integers(0, n_nodes, bad.sum())
    node1[rng.random(n + nd) < 0.02] = -1
    octave1 = rng.integers(0, 8, n + nd).astype(np.int32)
    angle1 = rng.uniform(0, 360, n + nd).astype(np.float32)
    angle2 = np.concatenate([angle1[:n] - 11.0 + rng.normal(0, 2.0, n), rng.uniform(0, 360, nd)]).astype(np.float32)
    out_rot = rng.random(n) < 0.05
    angle2[:n][out_rot] = rng.uniform(0, 360, out_rot.sum()).astype(np.float32)
    has1 = (rng.random(n + nd) < 0.2).astype(np.uint8); has2 = (rng.random(n + nd) < 0.2).astype(np.uint8)
    st1 = (rng.random(n + nd) < stereo_frac).astype(np.uint8); st2 = (rng.random(n + nd) < stereo_frac).astype(np.uint8)
    perm = rng.permutation(n + nd)                                  # keyframe 2 stores its keypoints in another order
    inv = np.empty_like(perm); inv[perm] = np.arange(n + nd)
    epipole = t21 / np.linalg.norm(t21)
    return dict(desc_1=d1, bearing_1=b1, octave_1=octave1, angle_1=angle1, has_lm_1=has1, is_stereo_1=st1, bow_node_1=node1,
                desc_2=d2[perm], bearing_2=b2[perm], angle_2=angle2[perm], has_lm_2=has2[perm], is_stereo_2=st2[perm], bow_node_2=node2[perm],
                E_12=E12, epipole_in_2=epipole, truth_idx_2_of_1=np.concatenate([inv[:n], -np.ones(nd, np.int64)]).astype(np.int32))
