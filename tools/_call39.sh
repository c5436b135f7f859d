exec 2>&1
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python bench.py --steps 200 --warmup 5 --quick > gpurun_out/q.json 2> gpurun_out/q.log
python - <<PY
import json
d=json.load(open("gpurun_out/q.json")); print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["resolve_kernels_ms"])
PY
done
