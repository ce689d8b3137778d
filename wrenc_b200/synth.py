"""Seeded synthetic I420 frames: smooth gradients + low-frequency sinusoids, oriented edges /
rectangles (exercise angular modes), band-limited noise (exercise the trellis), chroma correlated with luma
(exercise CCLM); a per-frame translation makes frames differ.  numpy only; no file or network access."""
import numpy as np


def synth_frame(width: int, height: int, seed: int = 0xB2000000, frame: int = 0):
    """Return (y, cb, cr) uint8 planes of one 4:2:0 frame."""
    assert width % 2 == 0 and height % 2 == 0
    rng = np.random.default_rng([seed & 0xFFFFFFFF, frame])
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    xx = xx + 3.0 * frame
    yy = yy + 1.0 * frame
    img = 110.0 + 0.08 * xx + 0.05 * yy
    for _ in range(4):
        fx, fy = rng.uniform(0.004, 0.05, 2)
        ph = rng.uniform(0, 6.28)
        img += rng.uniform(8, 25) * np.sin(fx * xx + fy * yy + ph)
    # oriented edges
    for _ in range(6):
        ang = rng.uniform(0, np.pi)
        off = rng.uniform(0, max(width, height))
        period = rng.uniform(40, 160)
        d = np.cos(ang) * xx + np.sin(ang) * yy + off
        img += rng.uniform(10, 35) * (np.floor(d / period) % 2)
    # rectangles
    for _ in range(max(2, (width * height) // 60000)):
        x0 = int(rng.integers(0, width)) - 3 * frame
        y0 = int(rng.integers(0, height))
        w = int(rng.integers(8, 96))
        h = int(rng.integers(8, 96))
        xs = slice(max(0, x0), max(0, min(width, x0 + w)))
        ys = slice(max(0, y0), max(0, min(height, y0 + h)))
        img[ys, xs] += rng.uniform(-50, 50)
    # band-limited noise, sigma ~ 6
    noise = rng.normal(0.0, 1.0, (height, width)).astype(np.float32)
    noise = (noise + np.roll(noise, 1, 0) + np.roll(noise, 1, 1) + np.roll(noise, (1, 1), (0, 1))) * 0.5
    img += 6.0 * noise
    y = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    ds = img.reshape(height // 2, 2, width // 2, 2).mean(axis=(1, 3))
    cnoise = rng.normal(0.0, 2.0, (2, height // 2, width // 2)).astype(np.float32)
    cb = 128.0 + 0.45 * (ds - 128.0) * np.sin(0.01 * xx[::2, ::2]) + 12.0 * np.sin(0.02 * yy[::2, ::2]) + cnoise[0]
    cr = 128.0 - 0.35 * (ds - 128.0) + 10.0 * np.cos(0.015 * xx[::2, ::2]) + cnoise[1]
    cb = np.clip(np.rint(cb), 0, 255).astype(np.uint8)
    cr = np.clip(np.rint(cr), 0, 255).astype(np.uint8)
    return y, cb, cr


def random_frame(width: int, height: int, seed: int):
    """Uniform random planes: a stress input (huge residuals) for parity tests."""
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 256, (height, width), dtype=np.uint8),
            rng.integers(0, 256, (height // 2, width // 2), dtype=np.uint8),
            rng.integers(0, 256, (height // 2, width // 2), dtype=np.uint8))
