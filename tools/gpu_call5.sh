mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.txt
timeout 300 python bench.py --workload generate > gpurun_out/bench_gen.json 2> gpurun_out/bench_gen.err; echo "gen rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_gen.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['cpu_baseline'])"
tail -3 gpurun_out/bench_gen.err
timeout 300 python bench.py --workload generate --trunc 0.95 --no-cpu-baseline > gpurun_out/bench_gen_trunc.json 2> gpurun_out/bench_gen_trunc.err; echo "gen trunc rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_gen_trunc.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
