#!/bin/bash
# usage: scripts/run_variants.sh gpurun_out/var_*.so   — bench line value per kernel-layout variant
for so in "$@"; do
  v=$(FB200_LIB=$PWD/$so timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('%.0f model-steps/s  %.1f ms  e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))" 2>&1 | tail -1)
  echo "$so: $v"
done
