"""tcgen05 TF32x3 GEMM path (csrc/conv_tc.cu) validated on the device against the exact-fp32 CUDA-core conv
(csrc/conv.cu) and against the CPU oracle, then end to end through the HMR forward."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

SHAPES = [  # M, Cin, Cout
    (3136, 64, 256), (3136, 256, 64), (784, 512, 128), (196, 1024, 256), (49, 2048, 512), (49, 512, 2048), (2 * 196, 256, 1024), (130, 64, 64)]


@pytest.fixture(scope='module')
def L():
    from dynaboa_b200 import _lib
    lib = _lib.load()
    yield _lib
    lib.dboa_set_tensor_core_conv(0)


def run_tc(L, mode, M, Cin, Cout, x, w, ws):
    L.load().dboa_set_tensor_core_conv(mode)
    y = torch.full((M, Cout), float('nan'), device='cuda')
    st = L.load().dboa_conv1x1_tc_fwd(L.ptr(x), L.ptr(w), L.ptr(y), M, Cin, Cout, L.ptr(ws), ws.numel(), L.stream())
    torch.cuda.synchronize()
    return st, y


def test_descriptor_convention_and_accuracy(L):
    """Reports which LBO/SBO role assignment the hardware accepts (mode 1 = as documented by CUTLASS)."""
    torch.manual_seed(0)
    ws = torch.empty(8 << 20, device='cuda')
    results = {}
    for mode in (1, 2):
        errs = []
        for M, Cin, Cout in SHAPES:
            x = torch.randn(M, Cin, device='cuda')
            w = torch.randn(Cout, Cin, device='cuda') / Cin ** 0.5
            st, y = run_tc(L, mode, M, Cin, Cout, x, w, ws)
            assert st == 0, (mode, M, Cin, Cout, st)
            ref = (x.double() @ w.double().t())
            errs.append(((y.double() - ref).abs().max() / ref.abs().max()).item())
        results[mode] = max(errs)
        print(f'tcgen05 mode {mode}: max rel err over shapes = {results[mode]:.3e}  per-shape {[f"{e:.1e}" for e in errs]}')
    assert results[1] < 5e-6, f'documented descriptor convention failed: {results}'


def test_matches_fp32_cuda_core_path(L):
    torch.manual_seed(1)
    ws = torch.empty(8 << 20, device='cuda')
    for M, Cin, Cout in SHAPES:
        x = torch.randn(M, Cin, device='cuda')
        w = torch.randn(Cout, Cin, device='cuda') / Cin ** 0.5
        st, y = run_tc(L, 1, M, Cin, Cout, x, w, ws)
        assert st == 0
        y32 = torch.empty(M, Cout, device='cuda')
        L.call('dboa_conv2d_fwd', L.ptr(x), L.ptr(w), L.ptr(y32), 1, M, 1, Cin, Cout, 1, 1, 0, Cin, L.ptr(ws), ws.numel(), L.stream())
        assert rel_err(y, y32) < 5e-6, (M, Cin, Cout)


def test_hmr_forward_with_tensor_core_convs(L, golden):
    from dynaboa_b200 import synthetic
    from dynaboa_b200.hmr import hmr
    from oracle import hmr_ref
    gd = golden('hmr_forward')
    m = hmr(synthetic.make_mean_params()).cuda()
    m.load_state_dict(hmr_ref.strip_prefix(synthetic.make_basemodel()['model']), strict=True)
    m.eval()
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(24)).cuda()
    L.load().dboa_set_tensor_core_conv(1)
    try:
        with torch.no_grad():
            rot, shape, cam, feats = m(x, need_feature=True)
    finally:
        L.load().dboa_set_tensor_core_conv(0)
    assert rel_err(rot, gd['rotmat']) < 1e-4 and rel_err(shape, gd['shape']) < 1e-4 and rel_err(cam, gd['cam']) < 1e-4
    for i in range(5, 15):
        assert rel_err(feats[i], gd[f'feat{i}']) < 1e-4, i
