mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.txt
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --shape h36m --no-cpu-baseline > gpurun_out/bench_h36m.json 2> gpurun_out/bench_h36m.err; echo "h36m rc=$?"
cut -c1-200 gpurun_out/bench_h36m.json
timeout 300 python bench.py --batch 32 --shape ntu60 --no-cpu-baseline > gpurun_out/bench_ntu60_b32.json 2> gpurun_out/bench_ntu60_b32.err; echo "b32 rc=$?"
cut -c1-200 gpurun_out/bench_ntu60_b32.json
