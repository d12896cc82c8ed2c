mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_trainer_gpu.py -m gpu -x -q > gpurun_out/pytest_quick.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_quick.txt
timeout 400 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json")); r=d["roofline"]
print(d["value"], d["ms_per_step"], r["traffic"])
for k,v in sorted(r["families"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:7]: print(k, round(v["ms_per_step"],3), v["launches_per_step"], v["gbs"] and round(v["gbs"]))
PY
tail -3 gpurun_out/bench_n1.err
