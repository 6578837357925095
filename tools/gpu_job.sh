mkdir -p gpurun_out
nproc
for B in 0 1; do
timeout 600 python bench.py --steps 12 --warmup 3 --streams 4 --blocking-sync $B --no-cpu-baseline > gpurun_out/bench_b$B.json 2> gpurun_out/bench_b$B.err; echo "bench blocking=$B rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_b$B.json'));print(d['value'],d['ms_per_step'],d['single_stream_ms_per_proof'],d['e2e']['value'],d['blocking_sync'])"; tail -2 gpurun_out/bench_b$B.err
done
