"""Aggregate the warp-stall samples of an `ncu --import-source on` report per CUDA source line, from the report alone:
  ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_lines_cuda.py src.csv [top]
(the export lists, per source file, every CUDA line with the samples of the SASS attributed to it; inlined device functions
appear under the file that defines them)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
secs, i = [], 0
while i < len(rows):
	r = rows[i]
	if len(r) >= 2 and r[0] == "File Path":
		secs.append([r[1], rows[i + 2], []])
		i += 3
		continue
	if secs:
		secs[-1][2].append(r)
	i += 1
tot, agg, stall, per_file = 0, collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
for f, h, d in secs:
	iS, iL = h.index("# Samples"), h.index("Line No")
	st = [k for k, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
	for r in d:
		if len(r) <= iS or not r[iL].strip():
			continue
		try:
			n = int(r[iS])
		except ValueError:
			continue
		key = (f.split("/")[-1], int(r[iL]))
		agg[key] += n; tot += n; per_file[key[0]] += n
		for k in st:
			try:
				v = int(r[k])
			except ValueError:
				v = 0
			if v:
				stall[key][h[k][6:]] += v
print("total samples", tot)
for f, c in per_file.most_common():
	print("%5.1f%% %s" % (100.0 * c / tot, f))
print()
for (f, l), c in agg.most_common(top):
	print("%5.1f%% %s:%d  [%s]" % (100.0 * c / tot, f, l, ", ".join("%s %d" % kv for kv in stall[(f, l)].most_common(3))))
