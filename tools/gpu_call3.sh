mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.txt
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --workload generate --no-cpu-baseline > gpurun_out/bench_gen.json 2> gpurun_out/bench_gen.err; echo "gen rc=$?"
cut -c1-200 gpurun_out/bench_gen.json; tail -3 gpurun_out/bench_gen.err
