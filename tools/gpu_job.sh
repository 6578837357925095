mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench default rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['ms_per_step'],d['single_stream_ms_per_proof'],d['e2e']['value'],d['roofline']['frac'],d['kernels'],d['cpu_baseline']['value'])"; tail -2 gpurun_out/bench.err
timeout 300 python tools/profile_proof.py 19 64 3 2>&1 | tail -1 | cut -c1-120
timeout 300 python tools/profile_proof.py 21 256 3 2>&1 | tail -1
timeout 300 python tools/profile_proof.py 23 1024 2 2>&1 | tail -1
