mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_dist_gpu.py -x -q > gpurun_out/pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -5 gpurun_out/pytest_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --msm-split --no-cpu-baseline > gpurun_out/bench_split_n2.json 2> gpurun_out/bench_split_n2.err; echo "split bench rc=$?"; tail -1 gpurun_out/bench_split_n2.json | cut -c1-600; tail -3 gpurun_out/bench_split_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "weak bench rc=$?"; tail -1 gpurun_out/bench_n2.json | cut -c1-400
