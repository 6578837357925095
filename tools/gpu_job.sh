mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
ZK_PHASE_TRACE=1 timeout 300 python tools/profile_proof.py 19 64 1 > gpurun_out/phase.log 2>&1; echo "phase rc=$?"; grep -E "phase|proof_wall" gpurun_out/phase.log | tail -12
timeout 300 python tools/profile_proof.py 19 64 5 2>&1 | tail -1
timeout 300 python tools/profile_proof.py 17 26 5 2>&1 | tail -1
timeout 300 python tools/profile_proof.py 21 256 3 2>&1 | tail -1
