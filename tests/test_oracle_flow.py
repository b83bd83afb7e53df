"""CPU: the torch-op flow used as the timed CPU arm (oracle/flow.py) equals the index-arithmetic oracle, and the
metric restatements behave."""
import numpy as np
import torch

from oracle import flow, metrics
from oracle import siren as osiren
from oracle import tiling as otiling
from oracle.synth import synth_image


def test_torch_flow_equals_numpy_oracle():
    torch.set_num_threads(2)
    sd = osiren.synth_state_dict(12, mod_bias_shift=0.5)
    img = synth_image(41, 80, 96)
    got = flow.reconstruct_slice(sd, torch.from_numpy(img)).numpy()
    patches, info = otiling.image_to_patches(img[None], 32, 16)
    kept, black, shape = otiling.filter_and_remember_black_patches(patches)
    assert len(black) > 0
    out = osiren.model_forward(sd, torch.from_numpy(kept)).numpy()
    want = otiling.patches_to_image_weighted_average(otiling.reintegrate_black_patches(out, black, shape), info, 24, 16)[0]
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
    p2, gs = flow.extract_patches(torch.from_numpy(img), 32, 16)
    assert gs == info[0] and np.array_equal(p2.numpy(), patches)


def test_metrics_restatement_properties():
    rs = np.random.RandomState(0)
    a = rs.uniform(size=(64, 64)).astype(np.float32)
    b = (a + rs.normal(scale=0.05, size=a.shape)).astype(np.float32)
    assert metrics.ssim(a, a) == 1.0
    assert metrics.nrmse(a, a) == 0.0
    assert 0.5 < metrics.ssim(a, b) < 1.0
    r = metrics.data_range(a, b)
    assert abs(metrics.psnr(a, b) - 10 * np.log10(r * r / np.mean((a.astype(np.float64) - b) ** 2))) < 1e-9
    assert metrics.psnr(a, b) > metrics.psnr(a, (a + 0.2).astype(np.float32))


def test_ssim_two_independent_formulations_and_closed_forms():
    """oracle/metrics.py is unpinned against scikit-image itself (absent here); what CAN be pinned: the filter-based
    restatement against a direct per-window evaluation that shares no code with it, and closed forms."""
    import numpy as np

    from oracle import metrics

    rs = np.random.RandomState(3)
    for shape in ((64, 80), (33, 100), (7, 9)):
        a = rs.uniform(0, 1, size=shape).astype(np.float32)
        b = (a + 0.1 * rs.normal(size=shape)).astype(np.float32)
        assert abs(metrics.ssim(a, b) - metrics.ssim_direct(a, b)) < 1e-12
        assert abs(metrics.ssim_direct(a, a) - 1.0) < 1e-15
    # y = x + c: contrast/structure terms cancel exactly, SSIM = mean over windows of the luminance term
    x = rs.uniform(0.2, 0.8, size=(40, 40))
    c = 0.1
    y = x + c
    from numpy.lib.stride_tricks import sliding_window_view

    mx = sliding_window_view(x, (7, 7)).mean(axis=(-1, -2))
    r = max(x.max(), y.max()) - min(x.min(), y.min())
    c1 = (0.01 * r) ** 2
    want = ((2 * mx * (mx + c) + c1) / (mx ** 2 + (mx + c) ** 2 + c1)).mean()
    assert abs(metrics.ssim(x, y) - want) < 1e-12
    # PSNR / NRMSE closed forms: y = x + c  =>  mse = c^2
    assert abs(metrics.psnr(x, y) - 10 * np.log10(r ** 2 / c ** 2)) < 1e-9
    assert abs(metrics.nrmse(x, y) - c / np.sqrt(np.mean(x ** 2))) < 1e-12
