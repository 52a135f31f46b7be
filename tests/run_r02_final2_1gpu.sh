#!/bin/bash
# round 2, final single-GPU evidence (session 3 code): GPU test suite, smoke, bench lines (both arms, all workloads)
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_fullsize.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_box.txt; nproc >> gpurun_out/r02_box.txt
( time python -m pytest tests/ -m gpu -q ) > gpurun_out/r02_pytest_gpu_final.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_final.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
python bench.py --bn eval --no-cpu-baseline --no-same-box > gpurun_out/r02_bench_1gpu_evalbn.json 2> /dev/null; echo "eval rc=$?"
python bench.py --workload baseline --no-cpu-baseline > gpurun_out/r02_bench_1gpu_baseline.json 2> /dev/null; echo "baseline rc=$?"
python bench.py --workload upperbound --no-cpu-baseline > gpurun_out/r02_bench_1gpu_upperbound.json 2> /dev/null; echo "upper rc=$?"
python bench.py --classes 4 --size 224 --no-cpu-baseline --no-same-box > gpurun_out/r02_bench_1gpu_acdc224.json 2> /dev/null; echo "acdc rc=$?"
python bench.py --classes 2 --size 224 --batch 96 --steps 20 --warmup 5 --no-cpu-baseline --no-same-box > gpurun_out/r02_bench_1gpu_lvsc224_b96.json 2> /dev/null; echo "lvsc rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_1gpu*.json")):
    try:
        d = json.load(open(f))
        print("%-48s %8.1f %s  %.3f ms (median %.3f)  e2e %s  roofline %.3f" % (
            f.split("/")[-1], d["value"], d["unit"], d["ms_per_step"], d.get("ms_per_step_median", 0),
            (d.get("e2e") or {}).get("value"), d["roofline"]["frac"]))
    except Exception as e:
        print(f, "failed", e)
PY
