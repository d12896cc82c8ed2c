mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tf32_gpu.py tests/test_trainer_gpu.py tests/test_generate_gpu.py -m gpu -x -q > gpurun_out/pytest_quick.txt 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_quick.txt
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json")); r=d["roofline"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k,v in sorted(r["families"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:6]: print(k, round(v["ms_per_step"],3), v["launches_per_step"], v["gbs"] and round(v["gbs"]))
for s in r["top_sites"][:14]: print(s["ms_per_step"], s["launches_per_step"], s["us_per_launch"], s["gbs"], s["site"])
PY
tail -3 gpurun_out/bench_n1.err
timeout 200 python tools/layer_bench.py --only D1 --batch 1024 > gpurun_out/layer_d1.txt 2>&1; cat gpurun_out/layer_d1.txt | tail -14
