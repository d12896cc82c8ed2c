mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.txt
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/bench_n1.json
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --shape h36m > gpurun_out/bench_h36m.json 2> gpurun_out/bench_h36m.err; echo "h36m rc=$?"; cut -c1-130 gpurun_out/bench_h36m.json
timeout 300 python bench.py --shape ntu60 --batch 32 --no-cpu-baseline > gpurun_out/bench_ntu60_b32.json 2> gpurun_out/bench_ntu60_b32.err; echo "b32 rc=$?"; cut -c1-130 gpurun_out/bench_ntu60_b32.json
timeout 300 python bench.py --workload generate > gpurun_out/bench_gen.json 2> gpurun_out/bench_gen.err; echo "gen rc=$?"; cut -c1-130 gpurun_out/bench_gen.json
timeout 300 python bench.py --workload generate --trunc 0.95 --trunc-cached --no-cpu-baseline > gpurun_out/bench_gen_trunc_cached.json 2> gpurun_out/bench_gen_trunc_cached.err; echo "gen trunc rc=$?"; cut -c1-130 gpurun_out/bench_gen_trunc_cached.json
timeout 200 python bench.py --workload generate --impl reference --steps 2 --warmup 1 > gpurun_out/bench_gen_ref.json 2> gpurun_out/bench_gen_ref.err; echo "gen ref rc=$?"
KGAN_NCU_RANGE=1 KGAN_NCU_STEPS=2 timeout 420 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/traffic.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
grep -c '^"' gpurun_out/traffic.csv
