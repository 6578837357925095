#!/bin/bash
# usage: tools_gpu_run.sh [tests] [bench] [ncu-launches] [ncu-full <kernel-regex>]
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests) timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log;;
smoke) timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log;;
bench) timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err;;
ncu-launches) timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1; echo "ncu-launches rc=$?";;
ncu-full=*) K=${what#ncu-full=}; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 2 -o gpurun_out/prof_$K -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo "ncu-full rc=$?";;
*) echo "running: $what"; timeout 1500 bash -c "$what";;
esac
done
