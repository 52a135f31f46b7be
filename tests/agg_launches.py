"""Aggregate an ncu per-launch CSV (gpu__time_duration.sum): python tests/agg_launches.py gpurun_out/launches.csv [steps]"""
import collections, csv, re, sys
fn = sys.argv[1]; steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = [r for r in csv.reader(open(fn, errors='ignore')) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    nm = re.sub(r'\(.*', '', r[ki]).replace('void pp::', '').replace('pp::', '')
    d[nm][0] += 1; d[nm][1] += float(r[vi].replace(',', ''))
tot = sum(v[1] for v in d.values())
print("%d launches, %.3f ms/step" % (len(rows), tot / 1e6 / steps))
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%9.1f us/step %6.1f x %5.1f%%  %s" % (v[1] / 1e3 / steps, v[0] / steps, 100 * v[1] / tot, k[:90]))
