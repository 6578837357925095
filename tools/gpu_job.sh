mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fixed_accumulate_kernel -s 8 -c 2 -o gpurun_out/prof_fixed_accumulate_v2 -f python tools/profile_proof.py 19 64 1 > gpurun_out/ncu2.log 2>&1; echo "ncu-full rc=$?"
