mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench default rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --streams 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1; echo "plain rc=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --streams 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1; echo "ncu-launches rc=$?"; wc -l gpurun_out/launches_bench.csv
