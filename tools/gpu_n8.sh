mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "n8 rc=$?"
cut -c1-260 gpurun_out/bench_n8.json; tail -5 gpurun_out/bench_n8.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload generate --no-cpu-baseline > gpurun_out/bench_gen_n8.json 2> gpurun_out/bench_gen_n8.err; echo "gen n8 rc=$?"
cut -c1-260 gpurun_out/bench_gen_n8.json
