#!/bin/bash
# Other shapes / variants on the same code (not the headline config): one short bench line each -> gpurun_out/variants.txt
mkdir -p gpurun_out
OUT=gpurun_out/variants.txt
: > $OUT
run() {
  echo "## bench.py $*" >> $OUT
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-profile-pass "$@" 2>> gpurun_out/variants.err \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f img/s  %.3f ms/step  clocks %s %s' % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])); print(d['config']['unet'], '| bn', d['config']['bn'], '| pairs/gpu', d['config']['pairs_per_gpu'])" >> $OUT
}
run
run --bn eval
run --unet-variant strided
run --unet-variant strided --output-stride 32
run --output-stride 32
run --size 224 --classes 4
run --size 224 --classes 2 --batch 96
run --batch 48
cat $OUT
