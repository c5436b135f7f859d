exec 2>&1
python tools/latency_probe.py 2>&1 | grep -E "latency|pipeline" | head -12
bash tools/gpu_round.sh r02final
