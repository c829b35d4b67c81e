"""Aggregates an ncu --csv launch list (gpu__time_duration.sum) by kernel: count, total ms, share.
agg_launches.py launches.csv [bench.log]: with the bench's own output (its JSON line holds gpu_launches of the timed steps) only
the LAST gpu_launches rows - the timed sweeps - are aggregated."""
import csv, re, sys, collections
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
if len(sys.argv) > 2:
	import json
	for l in open(sys.argv[2]):
		if l.startswith('{') and '"gpu_launches"' in l:
			n = int(json.loads(l)["gpu_launches"])
			print('timed region: last %d of %d launches' % (n, len(rows)))
			ours = [r for r in rows if 'at::' not in r['Kernel Name']]  # gpu_launches counts this library's kernels only
			rows = ours[-n:]
agg = collections.defaultdict(lambda: [0, 0.0])
def short(n):
	n = n.replace('(anonymous namespace)::', '').replace('void ', '')
	m = re.match(r'([A-Za-z_0-9]+)', n)
	base = m.group(1) if m else n[:40]
	if base == 'gemm_simt_kernel' or base.startswith('gemm_tc'):
		t = re.search(r'<(.*)>', n)
		return base + '<' + (t.group(1) if t else '') + '>'
	return base
for row in rows:
	v = float(row['Metric Value'].replace(',', ''))
	u = row['Metric Unit']
	v = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v * 1e3 if u in ('s', 'second') else v
	k = short(row['Kernel Name'])
	agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print('launches %d  total %.1f ms' % (len(rows), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
	print('%-80s n=%6d %10.2f ms %5.1f%%' % (k[:80], v[0], v[1], 100 * v[1] / tot))
