"""Top stalled SASS instructions of each launch in an ncu report (needs --import-source on / --set full).

    python scripts/ncu_hot.py gpurun_out/src_x.ncu-rep [top_n]
"""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--metrics',
                      'gpu__time_duration.sum,launch__grid_size,launch__block_size,launch__cluster_dim_x,launch__registers_per_thread'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader([l for l in raw.splitlines() if l.startswith('"')]))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('launch', d.get('ID'), {k: v for k, v in d.items() if k.startswith(('gpu__', 'launch__')) or k in ('Grid Size', 'Block Size')})
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = []
        blocks.append(cur)
    elif cur is not None and line.startswith('"'):
        cur.append(line)
for bi, b in enumerate(blocks):
    rr = list(csv.reader(b))
    h, body = rr[0], rr[1:]
    si = h.index('Warp Stall Sampling (All Samples)')
    stall_cols = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
    total = sum(int(r[si]) for r in body)
    agg = {h[i]: sum(int(r[i]) for r in body) for i in stall_cols}
    print(f'--- launch {bi}: {len(body)} SASS instructions, {total} samples; by reason:',
          ', '.join(f'{k[6:]}={v}' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    order = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:top]
    for i in sorted(order):
        r = body[i]
        why = max(stall_cols, key=lambda c: int(r[c]))
        print(f'  [{i:4d}] {int(r[si]):5d} {int(r[si]) / max(total, 1):6.1%} {h[why][6:]:12s} {r[1].strip()[:90]}')
