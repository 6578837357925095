mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "weak bench N=$N rc=$?"; tail -1 gpurun_out/bench_n$N.json | cut -c1-250; tail -3 gpurun_out/bench_n$N.err
