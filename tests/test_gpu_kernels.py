"""GPU tier, kernel level: the in-shared-memory FP64 FFT behind K2, through the C-ABI test hooks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B", [64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_packed_real_fft_roundtrip(gpu_lib, B):
    import torch
    L = gpu_lib.lib()
    n_ch = 5
    rng = np.random.default_rng(B)
    x = rng.standard_normal((n_ch, B))
    d_x = torch.from_numpy(x).cuda()
    d_spec = torch.zeros((n_ch, B, 2), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert L.dspb200_test_rfft(B, n_ch, d_x.data_ptr(), d_spec.data_ptr(), st) == 0, gpu_lib.last_error()
    torch.cuda.synchronize()
    spec = d_spec.cpu().numpy()
    spec = spec[..., 0] + 1j * spec[..., 1]
    want = np.fft.rfft(np.concatenate([x, np.zeros_like(x)], axis=1), axis=1)   # 2B-point transform of [x | 0]
    scale = np.max(np.abs(want))
    assert np.max(np.abs(spec[:, 1:] - want[:, 1:B])) < 1e-13 * scale
    assert np.max(np.abs(spec[:, 0].real - want[:, 0].real)) < 1e-13 * scale     # packed DC
    assert np.max(np.abs(spec[:, 0].imag - want[:, B].real)) < 1e-13 * scale     # packed Nyquist

    # inverse of a product spectrum = linear convolution of two B-blocks (length 2B-1)
    h = rng.standard_normal((n_ch, B))
    H = np.fft.rfft(np.concatenate([h, np.zeros_like(h)], axis=1), axis=1)
    Y = want * H
    packed = np.zeros((n_ch, B, 2))
    packed[:, 1:, 0] = Y[:, 1:B].real
    packed[:, 1:, 1] = Y[:, 1:B].imag
    packed[:, 0, 0] = Y[:, 0].real
    packed[:, 0, 1] = Y[:, B].real
    d_y = torch.from_numpy(packed).cuda()
    d_out = torch.zeros((n_ch, 2 * B), dtype=torch.float64, device="cuda")
    assert L.dspb200_test_irfft(B, n_ch, d_y.data_ptr(), d_out.data_ptr(), st) == 0, gpu_lib.last_error()
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    for c in range(n_ch):
        full = np.convolve(x[c], h[c])
        assert np.max(np.abs(out[c, :2 * B - 1] - full)) < 1e-12 * np.max(np.abs(full))
