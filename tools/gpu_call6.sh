mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['n_gpus'])"
tail -3 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload generate > gpurun_out/bench_gen_n2.json 2> gpurun_out/bench_gen_n2.err; echo "gen n2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_gen_n2.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['n_gpus'])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"
cut -c1-200 gpurun_out/bench_ref_n2.json
