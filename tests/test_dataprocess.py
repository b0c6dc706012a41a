"""Input side (N2): the oracle restatement of the reference's crop / resize / normalise / keypoint transform against the
fixture recorded from the reference's own code (tests/golden/dataprocess.npz, oracle/make_golden.py dataprocess), the host-side
filter composition of dynaboa_b200/dataprocess.py against the oracle (CPU), and the CUDA kernels against the oracle (GPU)."""
import numpy as np
import pytest
import torch


def test_oracle_reproduces_reference_fixture(golden):
    from oracle import dataprocess_ref as R
    gd = golden('dataprocess')
    for i, (center, scale) in enumerate(zip(gd['centers'], gd['scales'])):
        out = R.rgb_processing(gd['img'], list(center), float(scale))
        assert np.array_equal(out[:, ::4, ::4], gd['crops_sub'][i]), i
        assert np.array_equal(R.j2d_processing(gd['kp'], list(center), float(scale)), gd['kps'][i]), i


@pytest.mark.parametrize('n_in', [37, 224, 225, 300, 517])
def test_axis_matrix_is_the_resize_operator(n_in):
    """out = Wy . crop . Wx^T with the banded matrices equals the Gaussian-prefilter + zoom pipeline (float64 oracle)."""
    from dynaboa_b200.dataprocess import _axis_matrix
    from oracle import dataprocess_ref as R
    rng = np.random.default_rng(n_in)
    crop = rng.uniform(0, 255, size=(n_in, n_in + 13, 3))
    def dense(n):
        lo, w = _axis_matrix(n, 224)
        M = np.zeros((224, n))
        for o in range(224):
            T = min(w.shape[1], n - lo[o])
            M[o, lo[o]:lo[o] + T] = w[o, :T]
        return M
    Wy, Wx = dense(n_in), dense(n_in + 13)
    out = np.stack([Wy @ crop[:, :, c] @ Wx.T for c in range(3)], -1)
    ref = R.resize(crop, [224, 224])
    assert np.abs(out - ref).max() < 2e-4 * 255          # float32 weights


@pytest.mark.gpu
def test_gpu_crop_and_keypoints_match_the_oracle(golden):
    from dynaboa_b200 import dataprocess as D
    from oracle import dataprocess_ref as R
    gd = golden('dataprocess')
    img = torch.from_numpy(gd['img']).cuda()
    kp = torch.from_numpy(gd['kp']).float().cuda()
    for i, (center, scale) in enumerate(zip(gd['centers'], gd['scales'])):
        center, scale = list(center), float(scale)
        out = D.crop(img, center, scale)
        ref = R.rgb_processing(gd['img'], center, scale)
        assert out.shape == (3, 224, 224) and np.abs(out.cpu().numpy() - ref).max() < 2e-5, i
        assert np.array_equal(out.cpu().numpy()[:, ::4, ::4].round(4), gd['crops_sub'][i].round(4)) or np.abs(out.cpu().numpy()[:, ::4, ::4] - gd['crops_sub'][i]).max() < 2e-5
        k = D.j2d_processing(kp, center, scale)
        assert np.array_equal(k.cpu().numpy(), R.j2d_processing(gd['kp'].astype(np.float32).astype(np.float64), center, scale)), i
    # a large frame with the box partly outside it, uint8 input
    rng = np.random.default_rng(3)
    big = rng.integers(0, 256, size=(1080, 1920, 3), dtype=np.uint8)
    for center, scale in (([960.0, 540.0], 4.1), ([30.0, 1000.0], 2.2), ([1900.0, 20.0], 3.0)):
        out = D.crop(torch.from_numpy(big).cuda(), center, scale)
        ref = R.rgb_processing(big.astype(np.float32), center, scale)
        assert np.abs(out.cpu().numpy() - ref).max() < 3e-5, (center, scale)
