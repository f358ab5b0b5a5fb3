"""DEV TOOL: opcode histogram / stall summary of one kernel from an ncu report (source page), per tile and warp.
   python scripts/ncu_ops.py report.ncu-rep [tiles] [warps]"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
warps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
for k in ('gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct'):
    if k in hdr:
        print(k, vals[hdr.index(k)])
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
f = lambda r, k: float(r[ix[k]] or 0)
ops = collections.Counter()
tot = 0
for r in data:
    e = f(r, 'Instructions Executed')
    tot += e
    sp = r[ix['Source']].split()
    op = sp[1] if sp[0].startswith('@') else sp[0]
    ops[op] += e
print('warp instructions per tile per warp: %.1f' % (tot / tiles / warps))
for k, v in ops.most_common(28):
    print('  %-30s %.1f' % (k, v / tiles / warps))
samples = sum(f(r, '# Samples') for r in data)
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {k: sum(f(r, k) for r in data) for k in reasons}
print('stalls:', ', '.join('%s %.0f%%' % (k[6:], 100 * v / samples) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print('top sampled instructions:')
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:14]:
    print('  %5d  %s' % (f(r, '# Samples'), r[ix['Source']][:100]))
