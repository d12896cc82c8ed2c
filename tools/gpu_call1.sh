mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --shape h36m > gpurun_out/bench_h36m.json 2> gpurun_out/bench_h36m.err; echo "h36m rc=$?"
timeout 300 python bench.py --workload generate > gpurun_out/bench_gen.json 2> gpurun_out/bench_gen.err; echo "gen rc=$?"
KGAN_NCU_RANGE=1 timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/traffic.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/pytest_gpu.txt; cat gpurun_out/smoke.txt | tail -2; cut -c1-300 gpurun_out/bench_n1.json
