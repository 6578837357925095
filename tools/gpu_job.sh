timeout 300 python tools/profile_proof.py 19 64 5 2>&1 | tail -1
