mkdir -p gpurun_out
timeout 900 python tools/profile_proof.py 23 1024 2 2>&1 | tail -12
timeout 900 python tools/profile_proof.py 22 400 2 2>&1 | tail -3
