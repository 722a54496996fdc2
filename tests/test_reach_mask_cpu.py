"""The reach mask (csrc/sfgs_common.cuh::reach_mask) restated in numpy float32 and checked against the per-pixel test it
replaces: a (block, splat) pair may only be dropped if NO pixel of the 8x4 block satisfies the reference's blend
condition `power <= 0 and min(0.99, o * exp(power)) >= 1/255` (CR/forward.cu:400-417).  The mask only removes work, so
the property that matters is one-sided: every pixel the reference would blend lies in a block whose bit is set."""
import numpy as np
import pytest

f32 = np.float32


def reach_mask_np(mx, my, A, B, C, opac, tile_px, tile_py):
    """Same formulas and float32 operation order as the CUDA routine (approximate reciprocals replaced by exact ones)."""
    mx, my, A, B, C, opac = map(f32, (mx, my, A, B, C, opac))
    if not (opac >= f32(1.0 / 255.0)):
        return 0
    det = A * C - B * B
    if not (A > 0 and C > 0 and det > 0):
        return 0xFF
    xr, yr = mx - f32(tile_px), my - f32(tile_py)
    Xm = max(abs(xr), abs(xr - f32(15))); Ym = max(abs(yr), abs(yr - f32(15)))
    thr = f32(2.004) * np.log(opac * f32(255.0), dtype=f32) + f32(0.05) + f32(1e-5) * (A * Xm * Xm + f32(2) * abs(B) * Xm * Ym + C * Ym * Ym)
    invA, invC = f32(1) / A, f32(1) / C
    cols = [xr - f32((k >> 1) * 8 + (k & 1) * 7) for k in range(4)]
    rows = [yr - f32((j >> 1) * 4 + (j & 1) * 3) for j in range(8)]
    AX2 = [A * X * X for X in cols]; BX2 = [f32(2) * B * X for X in cols]; yc = [-B * X * invC for X in cols]
    CY2 = [C * Y * Y for Y in rows]; BY2 = [f32(2) * B * Y for Y in rows]; xc = [-B * Y * invA for Y in rows]
    mask = 0
    for by in range(4):
        yhi = yr - f32(4 * by); ylo = yhi - f32(3)
        for bx in range(2):
            xhi = xr - f32(8 * bx); xlo = xhi - f32(7)
            inside = ylo <= 0 <= yhi and xlo <= 0 <= xhi
            y0 = min(max(yc[2 * bx], ylo), yhi); y1 = min(max(yc[2 * bx + 1], ylo), yhi)
            q0 = (C * y0 + BX2[2 * bx]) * y0 + AX2[2 * bx]
            q1 = (C * y1 + BX2[2 * bx + 1]) * y1 + AX2[2 * bx + 1]
            x0 = min(max(xc[2 * by], xlo), xhi); x1 = min(max(xc[2 * by + 1], xlo), xhi)
            q2 = (A * x0 + BY2[2 * by]) * x0 + CY2[2 * by]
            q3 = (A * x1 + BY2[2 * by + 1]) * x1 + CY2[2 * by + 1]
            qmin = min(q0, q1, q2, q3)
            if inside or not (qmin > thr):
                mask |= 1 << (by * 2 + bx)
    return mask


def blended_blocks(mx, my, A, B, C, opac, tile_px, tile_py):
    """Bit per 8x4 block that holds at least one pixel the reference blends (its float32 expression, forward.cu:404-412)."""
    px = (np.arange(16, dtype=f32) + f32(tile_px))[None, :]
    py = (np.arange(16, dtype=f32) + f32(tile_py))[:, None]
    dx, dy = f32(mx) - px, f32(my) - py
    power = f32(-0.5) * (f32(A) * dx * dx + f32(C) * dy * dy) - f32(B) * dx * dy
    alpha = np.minimum(f32(0.99), f32(opac) * np.exp(power, dtype=f32))
    hit = (power <= 0) & (alpha >= f32(1.0 / 255.0))
    mask = 0
    for by in range(4):
        for bx in range(2):
            if hit[4 * by:4 * by + 4, 8 * bx:8 * bx + 8].any():
                mask |= 1 << (by * 2 + bx)
    return mask


def random_conic(rng, kind):
    """Conic (inverse 2D covariance) of a splat: round, elongated, or a needle hundreds of pixels long."""
    if kind == "round":
        s1 = s2 = rng.uniform(0.4, 12.0)
    elif kind == "elongated":
        s1, s2 = rng.uniform(0.4, 3.0), rng.uniform(3.0, 40.0)
    else:
        s1, s2 = rng.uniform(0.32, 0.6), rng.uniform(60.0, 600.0)
    th = rng.uniform(0, np.pi)
    c, s = np.cos(th), np.sin(th)
    cov = np.array([[c, -s], [s, c]]) @ np.diag([s1 * s1, s2 * s2]) @ np.array([[c, s], [-s, c]])
    inv = np.linalg.inv(cov)
    return f32(inv[0, 0]), f32(inv[0, 1]), f32(inv[1, 1])


@pytest.mark.parametrize("kind", ["round", "elongated", "needle"])
def test_reach_mask_never_drops_a_blended_pixel(kind):
    rng = np.random.default_rng({"round": 1, "elongated": 2, "needle": 3}[kind])
    dropped_pairs = kept_pairs = 0
    for _ in range(1500):
        A, B, C = random_conic(rng, kind)
        opac = float(rng.choice([rng.uniform(0.004, 0.02), rng.uniform(0.02, 1.0)]))
        tile_px, tile_py = 16 * int(rng.integers(0, 120)), 16 * int(rng.integers(0, 68))
        reach = 40.0 if kind != "needle" else 700.0           # centre inside or well outside the tile
        mx = tile_px + 8 + rng.uniform(-reach, reach)
        my = tile_py + 8 + rng.uniform(-reach, reach)
        m = reach_mask_np(mx, my, A, B, C, opac, tile_px, tile_py)
        need = blended_blocks(mx, my, A, B, C, opac, tile_px, tile_py)
        assert need & ~m == 0, (kind, mx, my, A, B, C, opac, tile_px, tile_py, bin(m), bin(need))
        dropped_pairs += bin(0xFF & ~m).count("1")
        kept_pairs += bin(m).count("1")
    assert dropped_pairs > 0          # the test frames do exercise the dropping branch


def test_reach_mask_special_cases():
    # opacity exactly 1/255 at the centre is blended by the reference (alpha == 1/255 is not < 1/255): kept
    o = float(np.float32(1.0 / 255.0))
    assert reach_mask_np(3.0, 2.0, 0.5, 0.0, 0.5, o, 0, 0) & 1          # centre (3, 2) lies in block 0
    assert reach_mask_np(3.0, 2.0, 0.5, 0.0, 0.5, np.nextafter(np.float32(o), np.float32(0)), 0, 0) == 0
    assert reach_mask_np(3.0, 2.0, 0.5, 0.0, 0.5, float("nan"), 0, 0) == 0
    # not a proper ellipse: everything is kept
    assert reach_mask_np(100.0, 100.0, 0.0, 0.0, 0.5, 0.9, 0, 0) == 0xFF
    assert reach_mask_np(100.0, 100.0, 0.5, 1.0, 0.5, 0.9, 0, 0) == 0xFF
