#!/bin/bash
# One gpurun call that does a whole measure-and-verify cycle on one B200 (about 3 minutes on the box): the gpu test
# suite, the full bench line, the ncu launch list of a short bench run and one ncu full capture of the scan kernel.
# Everything lands in gpurun_out/ under the tag given as $1 (default "round").
set -u
cd "$(dirname "$0")/.."
TAG=${1:-round}
mkdir -p gpurun_out
exec 2>&1
T0=$(date +%s)
el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
el start
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv,noheader
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
el "pytest rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.log
el "bench rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_slice|gather_kernel|chain_kernel|compact_kernel|init_header|shard_|small_|fec_kernel' -s 21 -c 42 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 12 --warmup 3 --quick --no-verify > gpurun_out/ncu_list_$TAG.log 2>&1
el "ncu list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:scan_slice -s 6 -c 1 -f -o gpurun_out/scan_$TAG python bench.py --steps 4 --warmup 3 --quick --no-verify > gpurun_out/ncu_full_$TAG.log 2>&1
el "ncu full rc=$?"
