"""Seeded synthetic frames of EuRoC / KITTI / TUM / 4K shape (datasets are unavailable offline).

Recipe: multi-octave value noise + random polygons/discs with random gray levels
(corners for FAST) + light blur + +-2 gray-level noise, quantised to uint8.  Shape counts are tuned so the
raw FAST@20 yield is a few times nfeatures (realistic for indoor/outdoor SLAM imagery), not the
noise-saturated regime.  numpy only; deterministic for a given (seed, shape).
"""
import numpy as np

BASE_SEED = 1234


def _smooth3(a):
    """separable [1 2 1]/4 blur, reflect border (float32)."""
    p = np.pad(a, 1, mode="reflect")
    a = (p[:, :-2] + 2 * p[:, 1:-1] + p[:, 2:]) * 0.25
    return (a[:-2] + 2 * a[1:-1] + a[2:]) * 0.25


def _value_noise(rng, h, w, cell):
    gh, gw = h // cell + 3, w // cell + 3
    g = rng.random((gh, gw), dtype=np.float32)
    ys = (np.arange(h, dtype=np.float32) + 0.5) / cell
    xs = (np.arange(w, dtype=np.float32) + 0.5) / cell
    y0 = ys.astype(np.int32)
    x0 = xs.astype(np.int32)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    fy = fy * fy * (3 - 2 * fy)
    fx = fx * fx * (3 - 2 * fx)
    a = g[y0][:, x0]
    b = g[y0][:, x0 + 1]
    c = g[y0 + 1][:, x0]
    d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def texture(h, w, seed, shapes_per_mpx=900.0):
    """float32 image in [0,255] before quantisation."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    img = np.zeros((h, w), np.float32)
    amp = 1.0
    tot = 0.0
    for cell in (64, 32, 16, 8, 4):
        img += amp * _value_noise(rng, h, w, cell)
        tot += amp
        amp *= 0.55
    img = 60.0 + 130.0 * img / tot
    n_shapes = max(8, int(shapes_per_mpx * h * w / 1e6))
    for _ in range(n_shapes):
        cx, cy = rng.integers(0, w), rng.integers(0, h)
        r = int(rng.integers(4, 28))
        x0, x1 = max(cx - r, 0), min(cx + r + 1, w)
        y0, y1 = max(cy - r, 0), min(cy + r + 1, h)
        if x1 - x0 < 2 or y1 - y0 < 2:
            continue
        yy, xx = np.ogrid[y0:y1, x0:x1]
        kind = rng.integers(0, 3)
        gray = float(rng.integers(10, 246))
        alpha = float(rng.uniform(0.55, 1.0))
        if kind == 0:                      # disc
            m = (xx - cx) ** 2 + (yy - cy) ** 2 <= (r * 0.7) ** 2
        else:                              # rotated rectangle (kind 1: square-ish, kind 2: elongated)
            th = float(rng.uniform(0, np.pi))
            ct, st = np.cos(th), np.sin(th)
            u = (xx - cx) * ct + (yy - cy) * st
            v = -(xx - cx) * st + (yy - cy) * ct
            hw_ = r * 0.68
            hh_ = hw_ * (1.0 if kind == 1 else float(rng.uniform(0.2, 0.6)))
            m = (np.abs(u) <= hw_) & (np.abs(v) <= hh_)
        sub = img[y0:y1, x0:x1]
        sub[m] = sub[m] * (1 - alpha) + gray * alpha
    img = _smooth3(img)
    img += rng.integers(-2, 3, size=(h, w)).astype(np.float32)
    return img


def frame(h, w, index=0, base_seed=BASE_SEED):
    """One uint8 frame; seed = base_seed + index."""
    return np.clip(np.rint(texture(h, w, base_seed + index)), 0, 255).astype(np.uint8)


def sequence(h, w, n, base_seed=BASE_SEED, canvas=2048, start=0, stop=None):
    """n frames produced as a slowly translating crop of one big texture (TUM-like sequence); start / stop select the frames
    [start, stop) of that n-frame sequence (a rank's contiguous block of a frame-partitioned sequence)."""
    big = np.clip(np.rint(texture(canvas, canvas, base_seed)), 0, 255).astype(np.uint8)
    stop = n if stop is None else stop
    out = np.empty((stop - start, h, w), np.uint8)
    span_x, span_y = canvas - w, canvas - h
    for i in range(start, stop):
        t = i / max(n - 1, 1)
        ox = int(round((0.5 + 0.5 * np.sin(2 * np.pi * t)) * span_x))
        oy = int(round((0.5 + 0.5 * np.cos(2 * np.pi * 0.5 * t)) * span_y))
        out[i - start] = big[oy:oy + h, ox:ox + w]
    return out


def stereo_pair(h, w, index=0, base_seed=BASE_SEED, dmin=4, dmax=64, band=47):
    """Rectified pair: right[y, x] = left[y, x + d(y)] with a piece-wise constant disparity per
    horizontal band, plus independent +-1 sensor noise on the right image."""
    rng = np.random.default_rng(np.random.PCG64(base_seed + 7919 * (index + 1)))
    wide = np.clip(np.rint(texture(h, w + dmax, base_seed + index)), 0, 255).astype(np.uint8)
    left = np.ascontiguousarray(wide[:, :w])
    right = np.empty_like(left)
    for y0 in range(0, h, band):
        d = int(rng.integers(dmin, dmax + 1))
        # a point at left column u appears at right column u - d
        y1 = min(y0 + band, h)
        src = np.arange(w) + d
        right[y0:y1] = wide[y0:y1][:, src]
    noise = rng.integers(-1, 2, size=right.shape)
    right = np.clip(right.astype(np.int32) + noise, 0, 255).astype(np.uint8)
    return left, right


def stereo_sequence(h, w, n, base_seed=BASE_SEED, dmin=4, dmax=64, band=47, step=5):
    """n rectified pairs cut from one wide texture (cheap enough for a 256-pair benchmark batch): the left image of
    pair i is the crop at column offset i*step; right[y, x] = left-texture[y, x + d_i(y)] with a per-pair, per-band
    constant disparity, plus +-1 noise.  -> (left[n,h,w], right[n,h,w]) uint8"""
    rng = np.random.default_rng(np.random.PCG64(base_seed + 104729))
    wide = np.clip(np.rint(texture(h, w + dmax + n * step, base_seed + 31)), 0, 255).astype(np.uint8)
    left = np.empty((n, h, w), np.uint8)
    right = np.empty((n, h, w), np.uint8)
    cols = np.arange(w)
    for i in range(n):
        off = i * step
        left[i] = wide[:, off:off + w]
        for y0 in range(0, h, band):
            d = int(rng.integers(dmin, dmax + 1))
            right[i, y0:y0 + band] = wide[y0:y0 + band][:, off + d + cols]
    noise = rng.integers(-1, 2, size=right.shape, dtype=np.int8)
    right = np.clip(right.astype(np.int16) + noise, 0, 255).astype(np.uint8)
    return left, right


def descriptor_db(n_db, n_query, seed=BASE_SEED, max_flips=40, dup_every=997):
    """kNN workload (BASELINE config 4 recipe): uniform random 256-bit database; half of the queries are
    database rows with k in [0, max_flips] random bit flips (planted neighbours), half uniform random;
    every dup_every-th database row duplicates its predecessor to exercise the lowest-index tie rule."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    db = rng.integers(0, 256, size=(n_db, 32), dtype=np.uint8)
    if dup_every and n_db > dup_every:
        db[dup_every::dup_every] = db[dup_every - 1::dup_every][: len(db[dup_every::dup_every])]
    q = rng.integers(0, 256, size=(n_query, 32), dtype=np.uint8)
    n_planted = n_query // 2
    src = rng.integers(0, n_db, size=n_planted)
    planted = db[src].copy()
    bits = np.unpackbits(planted, axis=1)
    flips = rng.integers(0, max_flips + 1, size=n_planted)
    pos = rng.integers(0, 256, size=(n_planted, max_flips))
    mask = np.arange(max_flips)[None, :] < flips[:, None]
    rows = np.repeat(np.arange(n_planted), max_flips)[mask.ravel()]
    np.bitwise_xor.at(bits, (rows, pos.ravel()[mask.ravel()]), 1)
    q[:n_planted] = np.packbits(bits, axis=1)
    return db, q
