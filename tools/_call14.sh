exec 2>&1
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 500 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.log
tail -2 gpurun_out/bench_r2j.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2j.json"))
print({k:d[k] for k in ("n_gpus","value","ms_per_step","hits_per_step")}, d["roofline"]["kernel_ms"], d["roofline"]["frac"])
print(d.get("e2e")); print(d.get("stream"))
PY
