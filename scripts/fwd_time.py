"""HMR forward (dboa_hmr_forward) timing, fused plan against the round-1 plan, L2 flushed between iterations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynaboa_b200 import _lib, hmr as hmr_mod, synthetic

lib = _lib.load()
if os.environ.get('FWD_KNOBS'):          # diagnostic library only (scripts/build_timeline.sh): wrong results, timing experiments
    lib.dboa_debug_set_fused_knobs(int(os.environ['FWD_KNOBS']))
m = hmr_mod.hmr(synthetic.make_mean_params()).cuda().eval()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for B in ((1, 9) if os.environ.get('FWD_KNOBS') else (1, 2, 9)):
    x = torch.randn(B, 3, 224, 224, device='cuda')
    tape = torch.empty(hmr_mod.tape_floats(B), device='cuda')
    for fused in ((1,) if os.environ.get('FWD_FUSED_ONLY') == '1' else (0, 1)):
        lib.dboa_set_fused_forward(fused)
        for flush_l2 in (True, False):
            ts = []
            for it in range(13):
                if flush_l2:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n0 = lib.dboa_launch_count()
                e0.record()
                hmr_mod.raw_forward(m.arena, m._buffers, x, tape=tape)
                e1.record()
                torch.cuda.synchronize()
                if it >= 3:
                    ts.append(e0.elapsed_time(e1))
            ts.sort()
            print(f'B={B} fused={fused} l2_flushed={flush_l2}: median {ts[len(ts)//2]*1000:.1f} us  min {ts[0]*1000:.1f} us  launches {lib.dboa_launch_count()-n0}')
lib.dboa_set_fused_forward(1)
