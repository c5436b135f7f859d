set -u
mkdir -p gpurun_out
exec 2>&1
T0=$(date +%s)
el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout 200 python tools/variant_check.py run --variants 'default' --iters 24 --no-best > gpurun_out/variants_r2f.log 2>&1
el "variants rc=$?"
cut -c1-700 gpurun_out/variants_r2f.log
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2f.log 2>&1
el "pytest rc=$? $(tail -1 gpurun_out/pytest_r2f.log)"
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.log
el "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2f.json"))
print({k:d[k] for k in ("n_gpus","value","ms_per_step","hits_per_step")}, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["resolve_kernels_ms"], d["verify"]["ok"])
for w in d.get("workloads",[]): print("  ", w["workload"][:40], w["format"], round(w["kernel_ms"],4), round(w["frac"],3), w["hits_per_step"], round(w["ms_per_step"],4))
PY
