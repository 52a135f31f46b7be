"""Summarise an ncu CSV with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch
(tests/run_ncu_round.sh -> gpurun_out/traffic.csv) into profiles/r01_conv_traffic.json, which bench.py reads for
`roofline.traffic` (DRAM bytes per launch of the dominant kernel).
usage: python tests/summarize_traffic.py gpurun_out/traffic.csv profiles/r01_conv_traffic.json"""
import collections, csv, json, re, sys

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src, errors="ignore")) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
per = collections.defaultdict(dict)
for r in rows:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("pp::", "")
    per[(r[ii], name)][r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for (_, name), m in per.items():
    f = fam[re.sub(r"<.*", "", name)]
    f[0] += 1
    f[1] += m.get("dram__bytes_read.sum", 0.0)
    f[2] += m.get("dram__bytes_write.sum", 0.0)
    f[3] += m.get("gpu__time_duration.sum", 0.0)
out = {k: {"launches": f[0], "dram_read_bytes_per_launch": f[1] / f[0], "dram_write_bytes_per_launch": f[2] / f[0],
           "avg_us": f[3] / f[0]} for k, f in fam.items()}
# the roofline's dominant kernel family = every forward / dgrad tcgen05 conv launch (bench.py's PROF_CONV family):
# the generic per-tap kernel, the shared-memory-resident rows kernel (round 2) and the narrow-layer halo kernel
dom = [k for k in out if re.fullmatch(r"conv3x3_(tc|rows_tc|halo_tc)_kernel", k)]
n_dom = sum(out[k]["launches"] for k in dom)
b_dom = sum(out[k]["launches"] * (out[k]["dram_read_bytes_per_launch"] + out[k]["dram_write_bytes_per_launch"]) for k in dom)
raw = sys.argv[3] if len(sys.argv) > 3 else src
res = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                 "(raw: %s), mean over the %d forward/dgrad conv launches (%s) of the profiled steps" % (
                     raw, n_dom, ", ".join(sorted(dom))),
       "conv3x3_tc_bytes_per_launch": b_dom / max(1, n_dom),
       "per_kernel": out}
json.dump(res, open(dst, "w"), indent=1)
for k, v in out.items():
    print("%-34s %4d launches  read %8.2f MB  write %7.2f MB  avg %7.1f us" % (
        k, v["launches"], v["dram_read_bytes_per_launch"] / 1e6, v["dram_write_bytes_per_launch"] / 1e6, v["avg_us"]))
