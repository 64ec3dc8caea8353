"""Summarise an `ncu --page source --csv` dump: stall mix, instruction mix and
the regions of the kernel by execution count (bin loops, producer, rest)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
name = rows[0][1]
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot_s = sum(int(r[col['# Samples']]) for r in body)
tot_i = sum(int(r[col['Instructions Executed']]) for r in body)
print(name)
print('warp-state samples: %d, warp instructions executed: %d' % (tot_s, tot_i))
mix = collections.Counter()
for r in body:
    for s in stalls:
        mix[s] += int(r[col[s]] or 0)
t = sum(mix.values())
print('stall mix (%):', ', '.join('%s %.1f' % (k[6:], 100. * v / t) for k, v in mix.most_common(9)))
op_i, op_s = collections.Counter(), collections.Counter()
for r in body:
    op = r[col['Source']].split()[0]
    if op.startswith('@'):
        op = r[col['Source']].split()[1]
    op = op.split('.')[0]
    op_i[op] += int(r[col['Instructions Executed']])
    op_s[op] += int(r[col['# Samples']])
print('instruction mix (% executed / % samples):',
      ', '.join('%s %.1f/%.1f' % (k, 100. * v / tot_i, 100. * op_s[k] / tot_s)
                for k, v in op_i.most_common(14)))
# regions by execution count
cnt = collections.Counter()
for r in body:
    cnt[int(r[col['Instructions Executed']])] += 1
print('regions (execution count per instruction: #instructions, % executed, % samples, stall mix):')
for c, n in sorted(cnt.items(), key=lambda kv: -kv[0] * kv[1])[:6]:
    rr = [r for r in body if int(r[col['Instructions Executed']]) == c]
    si = sum(int(r[col['# Samples']]) for r in rr)
    m = collections.Counter()
    for r in rr:
        for s in stalls:
            m[s] += int(r[col[s]] or 0)
    tt = max(1, sum(m.values()))
    print('  %12d x %4d instr: %5.1f %% executed, %5.1f %% samples | %s' % (
        c, n, 100. * c * n / tot_i, 100. * si / tot_s,
        ', '.join('%s %.0f' % (k[6:], 100. * v / tt) for k, v in m.most_common(5))))
