#!/bin/bash
# usage (on the GPU box, via gpurun): scripts/capture_profiles.sh <tag>  — the measurement set copied into profiles/ per kernel version:
# GPU tests, smoke, both bench arms, ncu launch list, one ncu --set full capture of the evolve kernel, phase cycles, other configs.
tag=${1:-vX}
out=gpurun_out
mkdir -p $out
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; tail -2 $out/${tag}_tests.log; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 400 $out/${tag}_bench.json
timeout 900 python bench.py --impl reference > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench.err; tail -c 300 $out/${tag}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fb200_evolve_kernel -s 5 -c 1 -f -o $out/${tag}_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
ls -la $out/${tag}_full.ncu-rep
FB200_LIB=$PWD/build/prof_default.so timeout 300 python scripts/phase_profile.py > $out/${tag}_phase_cycles.txt 2>&1; cat $out/${tag}_phase_cycles.txt
timeout 600 python scripts/bench_configs.py > $out/${tag}_other_configs.jsonl 2>&1; tail -5 $out/${tag}_other_configs.jsonl
if [ -n "$SMALL_NX" ]; then timeout 600 python scripts/small_nx_bench.py > $out/${tag}_small_nx.jsonl 2>&1; tail -12 $out/${tag}_small_nx.jsonl; fi
