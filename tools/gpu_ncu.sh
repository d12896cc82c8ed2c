mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tapconv_fwd_tma_k|tapconv_wgrad_tma_k" -c 12 -f -o gpurun_out/d1_full2 python tools/layer_bench.py --only D1 --reps 1 --batch 1024 > gpurun_out/ncu_d1.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/ncu_d1.log
ls -la gpurun_out/*.ncu-rep
