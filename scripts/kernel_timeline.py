"""Phase timeline inside the tcgen05 convolution kernel (diagnostic build).

    DBOA_TIMELINE=1 python -m dynaboa_b200.build --force      # rebuilds the library with %globaltimer stamps
    python scripts/kernel_timeline.py

For a few backbone shapes it launches the kernel in a short chain (so that programmatic dependent launch is in its
steady state), records thread 0's timestamps of the LAST launch for every CTA and prints, per phase, the median /
max duration over CTAs and the span from the first CTA's entry to the last CTA's exit.
Phases: 0 entry, 1 prologue done (barriers, TMEM), 2 weight prefetch issued, 3 dependency wait passed,
4 first k-block staged, 5 all MMAs issued, 6 MMAs complete, 7 accumulator read out, 8 cluster barrier,
9 split-K reduced and stored, 10 exit.
"""
import ctypes as C
import os
import sys

os.environ.setdefault('DBOA_CABI_PDL', '1')      # chains of the same kernel on fixed weights: PDL is safe for the stand-alone entry points

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from dynaboa_b200 import _lib as L  # noqa: E402

lib = L.load()
lib.dboa_set_tensor_core_conv(2)
lib.dboa_debug_set_timeline.restype = C.c_int
lib.dboa_debug_set_timeline.argtypes = [C.c_void_p]
NAMES = ['prologue', 'w-prefetch', 'dep-wait', 'stage k0', 'mma issue', 'mma done', 'tmem read', 'cl-barrier', 'reduce+st', 'exit']
SHAPES = [(56, 64, 64, 1, 1, 0), (56, 64, 64, 3, 1, 1), (28, 128, 512, 1, 1, 0), (14, 1024, 256, 1, 1, 0), (14, 256, 256, 3, 1, 1),
          (7, 512, 2048, 1, 1, 0), (7, 512, 512, 3, 1, 1)]
B = 1
buf = torch.zeros(4096 * 16, dtype=torch.int64, device='cuda')      # [cta][16] phase stamps; slots 40000.. = per-iteration cycle stamps of CTA 0
for mode in ('fwd', 'dgrad'):
    for (H, Cin, Cout, k, s, p) in SHAPES:
        Ho = (H + 2 * p - k) // s + 1
        K = k * k * Cin
        x = torch.randn(B, H, H, Cin, device='cuda')
        w = torch.randn(Cout, K, device='cuda') * 0.05
        y = torch.empty(B, Ho, Ho, Cout, device='cuda')
        dy = torch.randn(B, Ho, Ho, Cout, device='cuda')
        dx = torch.empty_like(x)
        a = (B, H, H, Cin, Cout, k, s, p, K)

        def run():
            if mode == 'fwd':
                L.call('dboa_conv2d_tc_fwd', L.ptr(x), L.ptr(w), L.ptr(y), *a, L.stream())
            else:
                L.call('dboa_conv2d_tc_dgrad', L.ptr(dy), L.ptr(w), L.ptr(dx), *a, 0, L.stream())
        lib.dboa_debug_set_timeline(None)
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        buf.zero_()
        lib.dboa_debug_set_timeline(C.c_void_p(buf.data_ptr()))
        for _ in range(4):
            run()
        torch.cuda.synchronize()
        clk = buf[40000:40000 + 96].view(12, 8).cpu()
        t = buf[:40000].view(-1, 16).cpu()
        t = t[t[:, 0] > 0][:, :11].double()
        ncta = t.shape[0]
        t0 = t[:, 0].min()
        d = t[:, 1:] - t[:, :-1]
        has_cl = (t[:, 8] > 0).all().item()
        print(f'--- {mode} H={H} Cin={Cin} Cout={Cout} k={k}: {ncta} CTAs, span first entry -> last exit {(t[:, 10].max() - t0) / 1e3:.2f} us, '
              f'entry skew {(t[:, 0].max() - t0) / 1e3:.2f} us, dep-wait passed at +{(t[:, 3].median() - t0) / 1e3:.2f} us (median)')
        if not has_cl:                      # no split-K: stamps 8, 9 are not written
            t[:, 8] = t[:, 7]
            t[:, 9] = t[:, 7]
            d = t[:, 1:] - t[:, :-1]
        print('    ' + '  '.join(f'{n}={d[:, i].median() / 1e3:.2f}/{d[:, i].max() / 1e3:.2f}' for i, n in enumerate(NAMES)) + '   (median/max us)')
        rows = [r for r in clk.tolist() if r[0] > 0]
        if rows:        # CTA 0: producer thread 0 [wait stage free, stash, fence+arrive, fetch issue] and MMA thread [wait full, issue 12 MMA + commit]
            print('    CTA0 cycles/iter producer[free,stash,arrive,fetch] total | mma[wait,issue]: ' + '  '.join(
                '[' + ','.join(str(r[j + 1] - r[j]) for j in range(4)) + f']{(rows[i + 1][0] - r[0]) if i + 1 < len(rows) else r[4] - r[0]}'
                + f'|[{r[6] - r[5]},{r[7] - r[6]}]' for i, r in enumerate(rows)))
