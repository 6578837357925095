mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
ZK_PHASE_TRACE=1 timeout 300 python tools/profile_proof.py 19 64 1 2>&1 | grep -E "phase" | tail -11
timeout 300 python tools/profile_proof.py 19 64 5 2>&1 | tail -1
