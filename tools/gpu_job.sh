timeout 300 python tools/microbench.py
