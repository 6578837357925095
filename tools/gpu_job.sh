mkdir -p gpurun_out
timeout 300 python tools/microbench.py > gpurun_out/microbench.json 2>&1; echo "microbench rc=$?"; cat gpurun_out/microbench.json
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
ZK_PHASE_TRACE=1 ZK_MSM_TRACE=1 timeout 300 python tools/profile_proof.py 19 64 1 > gpurun_out/phase.log 2>&1; echo "phase rc=$?"; tail -45 gpurun_out/phase.log
