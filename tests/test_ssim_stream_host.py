"""Host check of the arithmetic the row-streaming SSIM kernel uses (csrc/metrics.cu: ssim_stream_kernel): running 7-row
sums updated from the DIFFERENCE and SUM of the entering and leaving rows, window formula with the divisions by 49 / 48
cancelled.  Emulated here in fp32 with numpy (same recurrences, same segment restarts) and compared with the fp64
restatements in oracle/metrics.py -- so a wrong identity shows up without a GPU; the kernel itself is compared with
the same oracles in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import metrics as ometrics

f32 = np.float32


def ssim_stream_emulated(x: np.ndarray, y: np.ndarray, rows_per_seg: int) -> float:
    h, w = x.shape
    ny, nx = h - 6, w - 6
    mo, mp = f32(x.astype(np.float64).mean()), f32(y.astype(np.float64).mean())
    r = f32(max(x.max(), y.max())) - f32(min(x.min(), y.min()))
    c1, c2 = (f32(0.01) * r) ** 2, (f32(0.03) * r) ** 2
    C1, C2 = f32(2401.0) * c1, f32(48.0) * c2
    total = 0.0
    for y_a in range(0, ny, rows_per_seg):
        y_b = min(ny, y_a + rows_per_seg)
        vx = np.zeros(w, f32); vy = np.zeros(w, f32); vq = np.zeros(w, f32); vc = np.zeros(w, f32)
        for row in range(y_a, y_b + 6):
            a, b = x[row], y[row]
            if row - y_a >= 7:
                ao, bo = x[row - 7], y[row - 7]
            else:
                ao, bo = np.full(w, mo, f32), np.full(w, mp, f32)
            dx, sx = a - ao, (a + ao) - f32(2) * mo
            dy, sy = b - bo, (b + bo) - f32(2) * mp
            vx = vx + dx
            vy = vy + dy
            vq = dy * sy + (dx * sx + vq)
            vc = sx * dy + (dx * sy + vc)
            if row - y_a < 6:
                continue
            win = lambda v: np.lib.stride_tricks.sliding_window_view(v, 7).sum(-1, dtype=f32)
            Sx, Sy, Sq, Sc = win(vx), win(vy), win(vq), win(vc)
            Ux, Uy = Sx + f32(49) * mo, Sy + f32(49) * mp
            n1 = (Ux + Ux) * Uy + C1
            d1 = Uy * Uy + (Ux * Ux + C1)
            n2 = (Sx * Sy) * f32(-2.0 / 49.0) + (Sc + C2)
            d2 = (Sx * Sx + Sy * Sy) * f32(-1.0 / 49.0) + (Sq + C2)
            total += float(((n1 * n2) / (d1 * d2)).astype(np.float64).sum())
    return total / (ny * nx)


@pytest.mark.parametrize("hw,seg", [((64, 80), 16), ((40, 44), 34), ((120, 96), 57)])
def test_stream_recurrences_match_both_fp64_formulations(hw, seg):
    h, w = hw
    rs = np.random.RandomState(h + w)
    yy, xx = np.mgrid[0:h, 0:w]
    base = (0.5 + 0.4 * np.sin(xx / 7.0) * np.cos(yy / 5.0)).astype(f32)          # smooth structure + noise
    for scale, gain, off in ((0.01, 1.0, 0.0), (0.1, 0.5, -0.1), (0.3, 1.0, 0.0)):
        full = (base + rs.normal(scale=0.02, size=base.shape)).astype(f32)
        pred = ((full + rs.normal(scale=scale, size=base.shape)) * gain + off).astype(f32)
        got = ssim_stream_emulated(full, pred, seg)
        assert abs(got - ometrics.ssim(full, pred)) <= 2e-6
        assert abs(got - ometrics.ssim_direct(full, pred)) <= 2e-6
