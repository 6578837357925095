mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 3 --warmup 3 --msm-split --compressions 1024 --no-cpu-baseline > gpurun_out/bench_split_n${N}_k23.json 2> gpurun_out/bench_split_n${N}_k23.err; echo "split k23 N=$N rc=$?"; tail -1 gpurun_out/bench_split_n${N}_k23.json | cut -c1-330; tail -3 gpurun_out/bench_split_n${N}_k23.err
nvidia-smi --query-gpu=memory.used --format=csv | head -3
