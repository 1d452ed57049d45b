"""Seeded synthetic camera frames.

One generator shared by the tests, the golden-vector script and bench.py, so that the CUDA path,
the oracle and the host-built reference all see identical bytes.  Randomness comes from a
vectorised splitmix64 so frames do not depend on the numpy version.

Families
  noise  (F1)  every byte uniform 0..255 (worst case for label count and histograms)
  scene  (F2)  constant background colour, 1-3 filled rectangles / discs of a second colour,
               +-2 LSB luma noise, and a dark vertical band (what a line sensor looks at)
  grid   (F2)  every cell of an MxN grid painted its own colour with ~10 % distractor pixels
  blobs  (F2)  10-20 saturated red blobs of different sizes + isolated specks on a dark frame
  edge   (F3)  all-zero, all-255, blue-wrap (Y=U=255), grey ramp, metapixel checkerboard, 8 specks

Layouts
  "yuyv"     webcam sensors: bytes Y0 U Y1 V, row stride line_length >= 2*W
             (webcam/object_sensor/include/internal/cv_ball_detector_seqpass.hpp:251-284)
  "yuv422p"  ov7670 sensors: W x H luma plane, then at line_length*H a plane of interleaved
             chroma whose bytes are V U V U ... (ov7670/object_sensor/.../cv_ball_detector_seqpass.hpp:343-373)
"""
import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed, n, stream=0):
    """n uint64 values of the splitmix64 sequence started at (seed, stream)."""
    with np.errstate(over="ignore"):
        base = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) ^ (np.uint64(stream & 0xFFFFFFFF) * np.uint64(0xD1342543DE82EF95))
        z = base + _GOLDEN * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _rand_bytes(seed, n, stream):
    words = splitmix64(seed, (n + 7) // 8, stream)
    return words.view(np.uint8)[:n].copy()


def _randint(seed, stream, lo, hi, n=1):
    """n integers in [lo, hi) (hi > lo)."""
    r = splitmix64(seed, n, stream)
    return (lo + (r % np.uint64(hi - lo)).astype(np.int64)).astype(np.int64)


def pack(y, u, v, layout, line_length=None):
    """y: HxW, u/v: Hx(W/2) uint8 planes -> flat uint8 frame in the sensor's layout."""
    h, w = y.shape
    if layout == "yuyv":
        line = 2 * w if line_length is None else line_length
        out = np.zeros((h, line), dtype=np.uint8)
        out[:, 0:2 * w:4] = y[:, 0::2]
        out[:, 1:2 * w:4] = u
        out[:, 2:2 * w:4] = y[:, 1::2]
        out[:, 3:2 * w:4] = v
        return out.reshape(-1)
    if layout == "yuv422p":
        line = w if line_length is None else line_length
        out = np.zeros((2 * h, line), dtype=np.uint8)
        out[:h, :w] = y
        out[h:, 0:w:2] = v
        out[h:, 1:w:2] = u
        return out.reshape(-1)
    raise ValueError(layout)


def frame_bytes(w, h, layout, line_length=None):
    if layout == "yuyv":
        return h * (2 * w if line_length is None else line_length)
    return 2 * h * (w if line_length is None else line_length)


def planes_noise(seed, w, h):
    y = _rand_bytes(seed, w * h, 1).reshape(h, w)
    u = _rand_bytes(seed, w * h // 2, 2).reshape(h, w // 2)
    v = _rand_bytes(seed, w * h // 2, 3).reshape(h, w // 2)
    return y, u, v


def planes_scene(seed, w, h, band=True):
    p = _randint(seed, 10, 0, 256, 64)
    y = np.full((h, w), 40 + p[0] % 176, dtype=np.int32)
    u = np.full((h, w // 2), p[1], dtype=np.int32)
    v = np.full((h, w // 2), p[2], dtype=np.int32)
    nshapes = 1 + int(p[3] % 3)
    rows = np.arange(h)[:, None]
    cols = np.arange(w)[None, :]
    cols2 = np.arange(w // 2)[None, :] * 2
    for k in range(nshapes):
        q = p[8 + 8 * k: 16 + 8 * k]
        cy, cx = int(q[0]) * h // 256, int(q[1]) * w // 256
        ry, rx = 4 + int(q[2]) * h // 1024, 4 + int(q[3]) * w // 1024
        if q[7] & 1:
            m = (np.abs(rows - cy) <= ry) & (np.abs(cols - cx) <= rx)
            m2 = (np.abs(rows - cy) <= ry) & (np.abs(cols2 - cx) <= rx)
        else:
            m = (rows - cy) ** 2 * rx * rx + (cols - cx) ** 2 * ry * ry <= rx * rx * ry * ry
            m2 = (rows - cy) ** 2 * rx * rx + (cols2 - cx) ** 2 * ry * ry <= rx * rx * ry * ry
        y[m] = 16 + int(q[4]) % 224
        u[m2] = int(q[5])
        v[m2] = int(q[6])
    if band:
        bx = 8 + int(p[40]) * (w - 48) // 256
        bw = 6 + int(p[41]) % 28
        y[:, bx:bx + bw] = int(p[42]) % 24
        u[:, bx // 2:(bx + bw) // 2] = 128
        v[:, bx // 2:(bx + bw) // 2] = 128
    noise = (_rand_bytes(seed, w * h, 11).reshape(h, w) % 5).astype(np.int32) - 2
    y = np.clip(y + noise, 0, 255)
    return y.astype(np.uint8), u.astype(np.uint8), v.astype(np.uint8)


def planes_camera(seed, w, h, m=0, n=0):
    """A scene (or, with m x n given, a painted grid) as a real sensor delivers it: independent noise of +-2 LSB on the
    luma AND on both chroma channels of every sample."""
    if m and n:
        p = _randint(seed, 20, 0, 256, 3 * m * n)
        y = np.zeros((h, w), dtype=np.int32)
        u = np.zeros((h, w // 2), dtype=np.int32)
        v = np.zeros((h, w // 2), dtype=np.int32)
        hs, ws = h // m, w // n
        for i in range(m):
            for j in range(n):
                k = 3 * (i * n + j)
                y[i * hs:(i + 1) * hs, j * ws:(j + 1) * ws] = 32 + p[k] % 208
                u[i * hs:(i + 1) * hs, (j * ws) // 2:((j + 1) * ws) // 2] = p[k + 1]
                v[i * hs:(i + 1) * hs, (j * ws) // 2:((j + 1) * ws) // 2] = p[k + 2]
    else:
        y, u, v = (a.astype(np.int32) for a in planes_scene(seed, w, h))
    y = y + (_rand_bytes(seed, w * h, 31).reshape(h, w) % 5).astype(np.int32) - 2
    u = u + (_rand_bytes(seed, w * h // 2, 32).reshape(h, w // 2) % 5).astype(np.int32) - 2
    v = v + (_rand_bytes(seed, w * h // 2, 33).reshape(h, w // 2) % 5).astype(np.int32) - 2
    return (np.clip(y, 0, 255).astype(np.uint8), np.clip(u, 0, 255).astype(np.uint8), np.clip(v, 0, 255).astype(np.uint8))


def planes_grid(seed, w, h, m, n):
    """m rows x n columns of cells (the mxn sensor's widthM x heightN, names as the reference swaps them)."""
    p = _randint(seed, 20, 0, 256, 3 * m * n)
    y = np.zeros((h, w), dtype=np.uint8)
    u = np.zeros((h, w // 2), dtype=np.uint8)
    v = np.zeros((h, w // 2), dtype=np.uint8)
    hs, ws = h // m, w // n
    for i in range(m):
        for j in range(n):
            k = 3 * (i * n + j)
            y[i * hs:(i + 1) * hs, j * ws:(j + 1) * ws] = 32 + p[k] % 208
            u[i * hs:(i + 1) * hs, (j * ws) // 2:((j + 1) * ws) // 2] = p[k + 1]
            v[i * hs:(i + 1) * hs, (j * ws) // 2:((j + 1) * ws) // 2] = p[k + 2]
    d = _rand_bytes(seed, w * h, 21).reshape(h, w)
    dy = _rand_bytes(seed, w * h, 22).reshape(h, w)
    mask = d < 26
    y[mask] = dy[mask]
    d2 = _rand_bytes(seed, w * h // 2, 23).reshape(h, w // 2)
    du = _rand_bytes(seed, w * h // 2, 24).reshape(h, w // 2)
    dv = _rand_bytes(seed, w * h // 2, 25).reshape(h, w // 2)
    mask2 = d2 < 26
    u[mask2] = du[mask2]
    v[mask2] = dv[mask2]
    return y, u, v


def planes_blobs(seed, w, h):
    """Dark frame with 10-20 separate saturated red blobs of different sizes plus isolated 4x4 specks:
    what the object sensor ranks (>= 8 labels, distinct and equal sizes)."""
    p = _randint(seed, 30, 0, 1 << 16, 256)
    y = np.full((h, w), 24, dtype=np.uint8)
    u = np.full((h, w // 2), 128, dtype=np.uint8)
    v = np.full((h, w // 2), 128, dtype=np.uint8)
    nblobs = 10 + int(p[0] % 11)
    for k in range(nblobs):
        q = p[4 + 4 * k: 8 + 4 * k]
        bh = 4 * (1 + int(q[2]) % max(1, h // 24))
        bw = 4 * (1 + int(q[3]) % max(1, w // 32))
        r = 4 * (int(q[0]) % max(1, (h - bh) // 4))
        c = 4 * (int(q[1]) % max(1, (w - bw) // 4))
        y[r:r + bh, c:c + bw] = 120
        u[r:r + bh, c // 2:(c + bw) // 2] = 90
        v[r:r + bh, c // 2:(c + bw) // 2] = 240
    for k in range(12):
        q = p[128 + 2 * k: 130 + 2 * k]
        r = 4 * (int(q[0]) % (h // 4))
        c = 4 * (int(q[1]) % (w // 4))
        y[r:r + 4, c:c + 4] = 120
        u[r:r + 4, c // 2:c // 2 + 2] = 90
        v[r:r + 4, c // 2:c // 2 + 2] = 240
    return y, u, v


EDGE_CASES = ("zero", "full", "bluewrap", "greyramp", "checker", "specks", "red", "halves")


def planes_edge(name, w, h):
    y = np.zeros((h, w), dtype=np.uint8)
    u = np.full((h, w // 2), 128, dtype=np.uint8)
    v = np.full((h, w // 2), 128, dtype=np.uint8)
    if name == "zero":
        u[:] = 0
        v[:] = 0
    elif name == "full":
        y[:] = 255
        u[:] = 255
        v[:] = 255
    elif name == "bluewrap":       # 129*U - 17672 + 74*Y > 32767: the blue lane wraps negative
        y[:] = np.linspace(0, 255, w, dtype=np.uint8)[None, :]
        u[:] = np.linspace(200, 255, h, dtype=np.uint8)[:, None]
        v[:] = 100
    elif name == "greyramp":
        y[:] = (np.arange(w) * 255 // (w - 1)).astype(np.uint8)[None, :]
    elif name == "checker":        # 4x4 metapixel checkerboard, red-ish on / dark off
        on = (((np.arange(h)[:, None] // 4) + (np.arange(w)[None, :] // 4)) & 1) == 1
        y[:] = np.where(on, 120, 20)
        on2 = on[:, 0::2]
        u[:] = np.where(on2, 90, 128)
        v[:] = np.where(on2, 230, 128)
    elif name == "specks":         # dark frame with 8+ isolated bright red 4x4 specks and one blob
        y[:] = 16
        for k in range(10):
            r, c = 8 + 20 * (k % 5) * h // 120, 16 + 56 * (k // 5 + k % 3) * w // 320
            r, c = (r // 4) * 4, (c // 4) * 4
            y[r:r + 4, c:c + 4] = 130
            u[r:r + 4, c // 2:c // 2 + 2] = 90
            v[r:r + 4, c // 2:c // 2 + 2] = 240
        r0, c0 = (h // 2 // 4) * 4, (w // 2 // 4) * 4
        y[r0:r0 + 40, c0:c0 + 48] = 130
        u[r0:r0 + 40, c0 // 2:c0 // 2 + 24] = 90
        v[r0:r0 + 40, c0 // 2:c0 // 2 + 24] = 240
    elif name == "red":
        y[:] = 110
        u[:] = 90
        v[:] = 240
    elif name == "halves":         # left dark, right bright
        y[:, w // 2:] = 235
    else:
        raise ValueError(name)
    return y, u, v


def make_frame(family, seed, w, h, layout, line_length=None, **kw):
    """Flat uint8 frame.  family: 'noise' | 'scene' | 'grid' | 'blobs' | 'camera' | one of EDGE_CASES."""
    if family == "noise":
        planes = planes_noise(seed, w, h)
    elif family == "scene":
        planes = planes_scene(seed, w, h, band=kw.get("band", True))
    elif family == "grid":
        planes = planes_grid(seed, w, h, kw.get("m", 3), kw.get("n", 3))
    elif family == "blobs":
        planes = planes_blobs(seed, w, h)
    elif family == "camera":
        planes = planes_camera(seed, w, h, kw.get("m", 0), kw.get("n", 0))
    else:
        planes = planes_edge(family, w, h)
    return pack(planes[0], planes[1], planes[2], layout, line_length)


def make_batch(family, seeds, w, h, layout, out=None, **kw):
    """Stack of frames, shape (len(seeds), frame_bytes)."""
    nbytes = frame_bytes(w, h, layout, kw.get("line_length"))
    if out is None:
        out = np.empty((len(seeds), nbytes), dtype=np.uint8)
    for i, s in enumerate(seeds):
        out[i] = make_frame(family, int(s), w, h, layout, **kw)
    return out
