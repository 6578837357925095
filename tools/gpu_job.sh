mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fixed_accumulate_kernel -s 21 -c 1 -o gpurun_out/prof_fixed_accumulate_v2 -f python tools/profile_proof.py 19 64 1 > gpurun_out/ncu2.log 2>&1; echo "ncu-full rc=$?"
