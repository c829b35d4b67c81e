"""Aggregate the stall samples of an `ncu --page source --csv` export per CUDA source line (SASS rows are matched in order
with `nvdisasm -g` of a cubin built from the same source with the same flags). usage: ncu_lines.py src.csv kernel.cubin [lo hi]"""
import csv, re, subprocess, collections, sys
out = subprocess.run(["nvdisasm", "-g", sys.argv[2]], capture_output=True, text=True).stdout
lines, cur = [], ("?", 0)
for l in out.splitlines():
	m = re.search(r'//## File "([^"]+)", line (\d+)', l)
	if m:
		cur = (m.group(1).split('/')[-1], int(m.group(2)))
		continue
	if re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l):
		lines.append(cur)
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
assert len(data) == len(lines), (len(data), len(lines))
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in data)
if len(sys.argv) > 4:  # SASS rows of a source-line window, in order
	lo, hi = int(sys.argv[3]), int(sys.argv[4])
	inside = False
	for k, (r, l) in enumerate(zip(data, lines)):
		if l[0].endswith(".cu"):
			inside = lo <= l[1] <= hi
		if inside and int(r[iS]) >= 3:
			st = sorted(((int(r[i]) if r[i] else 0, hdr[i][6:]) for i in stall), reverse=True)[:2]
			print(k, r[iS], r[iEx], r[iSrc].strip()[:60], l, st)
else:
	agg, ex, st = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
	for r, l in zip(data, lines):
		agg[l] += int(r[iS]); ex[l] += int(r[iEx])
		for i in stall:
			v = int(r[i]) if r[i] else 0
			if v: st[l][hdr[i][6:]] += v
	print("total samples", tot)
	for l, c in agg.most_common(40):
		print("%5.1f%% %6d ex=%9d %s:%d  [%s]" % (100 * c / tot, c, ex[l], l[0], l[1], ", ".join("%s %d" % kv for kv in st[l].most_common(3))))
