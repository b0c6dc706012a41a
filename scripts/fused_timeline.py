"""Per-launch phase timeline of the fused forward (diagnostic build):

    DBOA_TIMELINE=1 python -m dynaboa_b200.build --force && python scripts/fused_timeline.py [B]

Stamps (thread 0 of each CTA, %globaltimer ns): 0 entry, 1 set-up done, 2 dependency wait passed, 3 first k-block staged,
4 producer loop done, 5 MMAs complete, 6 accumulators in shared memory (+ cluster barrier), 7 outputs and statistics stored,
8 exit.  Printed per launch: grid, start of the first CTA relative to the end of the previous launch, and the median over
CTAs of every phase duration."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynaboa_b200 import _lib, hmr as hmr_mod, synthetic  # noqa: E402

lib = _lib.load()
lib.dboa_debug_set_fused_timeline.restype = C.c_int
lib.dboa_debug_set_fused_timeline.argtypes = [C.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m = hmr_mod.hmr(synthetic.make_mean_params()).cuda().eval()
x = torch.randn(B, 3, 224, 224, device='cuda')
tape = torch.empty(hmr_mod.tape_floats(B), device='cuda')
for _ in range(5):
    hmr_mod.raw_forward(m.arena, m._buffers, x, tape=tape)
torch.cuda.synchronize()
buf = torch.zeros(128 * 256 * 16 + 128 * 16 * 8 + 128 * 32, dtype=torch.int64, device='cuda')
lib.dboa_debug_set_fused_timeline(C.c_void_p(buf.data_ptr()))
hmr_mod.raw_forward(m.arena, m._buffers, x, tape=tape)
torch.cuda.synchronize()
lib.dboa_debug_set_fused_timeline(None)
t = buf[:128 * 256 * 16].view(128, 256, 16).cpu()
ti = buf[128 * 256 * 16:128 * 256 * 16 + 128 * 16 * 8].view(128, 16, 8).cpu()
tm = buf[128 * 256 * 16 + 128 * 16 * 8:].view(128, 32).cpu()
names = ['setup', 'dep-wait', 'stage k0', 'k-loop', 'mma done', 'tmem+sync', 'store+stats', 'cl-sync']
prev_end = None
print(f'B={B}   launch  ctas |  gap(us) span(us) | ' + ' '.join(f'{n:>11s}' for n in names))
tot = 0.0
for l in range(128):
    rows = t[l][t[l][:, 0] > 0]
    if rows.numel() == 0:
        continue
    start, end = rows[:, 0].min().item(), rows[:, 8].max().item()
    d = (rows[:, 1:9] - rows[:, 0:8]).double() / 1000.0
    med = d.median(0).values
    gap = (start - prev_end) / 1000.0 if prev_end is not None else 0.0
    span = (end - start) / 1000.0
    tot += span
    ep = torch.stack([rows[:, 9] - rows[:, 5], rows[:, 6] - rows[:, 9], rows[:, 10] - rows[:, 6], rows[:, 11] - rows[:, 10],
                      rows[:, 12] - rows[:, 11], rows[:, 7] - rows[:, 12]], 1).double().median(0).values / 1000.0
    print(f'        {l:5d} {rows.shape[0]:5d} | {gap:8.2f} {span:8.2f} | ' + ' '.join(f'{v:11.2f}' for v in med.tolist())
          + '  || epilogue: tmem->smem %.2f sync %.2f rows+store %.2f fixed-point+shfl %.2f smem atomics+sync %.2f global atomics %.2f' % tuple(ep.tolist()))
    prev_end = end
print(f'sum of spans {tot:.1f} us')

print('per-k-block stamps of CTA 0 (us relative to its dependency wait): loop top | tiles landed | lo free (tensor-memory operand: loads + math done) | arrive | - (tensor-memory operand: operand stage free) || mma: lo-full | issued || tma issued;  with the tensor-memory operand even / odd k-blocks belong to transform group 0 / 1')
for l in (0, 1, 2, 3, 20, 28, 29, 30):
    base = t[l][0][2].item()
    if base == 0:
        continue
    print(f'launch {l}')
    for it in range(16):
        row = ti[l][it]
        if row.max().item() == 0:
            break
        print('   it %2d: ' % it + ' '.join('%8.2f' % ((v - base) / 1000.0) if v > 0 else '     -  ' for v in row.tolist()))

