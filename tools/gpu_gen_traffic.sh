mkdir -p gpurun_out
KGAN_NCU_RANGE=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/traffic_gen.csv python bench.py --workload generate --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_gen_ncu.log 2>&1; echo "ncu rc=$?"
grep -c '^"' gpurun_out/traffic_gen.csv
