mkdir -p gpurun_out
for S in 4 6; do
timeout 600 python bench.py --steps 12 --warmup 3 --streams $S --no-cpu-baseline > gpurun_out/bench_s$S.json 2> gpurun_out/bench_s$S.err; echo "bench S=$S rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_s$S.json'));print(d['value'],d['ms_per_step'],d['single_stream_ms_per_proof'],d['e2e']['value'],d['roofline']['frac'],d['gpu_launches'])"; tail -2 gpurun_out/bench_s$S.err
done
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench default rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
