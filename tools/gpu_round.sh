#!/bin/bash
# One gpurun call that does a whole measure-and-verify cycle (about 2 minutes on the box): A/B the scan-kernel
# variants built by `python tools/variant_check.py build` (parity against the oracle and the golden vectors +
# headline timing per variant), then bench.py, the gpu test suite, the ncu launch list and one ncu full capture
# with the fastest parity-green one.  Everything lands in gpurun_out/.   VARIANTS=a,b,c bash tools/gpu_round.sh
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
exec 2>&1
T0=$(date +%s)
el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
el start
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv,noheader
timeout 170 python tools/variant_check.py run --variants ${VARIANTS:-default} --iters 24 > gpurun_out/variants.log 2>&1
el "variants rc=$?"
BESTFULL=$(cat gpurun_out/best_variant.txt 2>/dev/null || echo base)
BEST=${BESTFULL%@*}
if [ "$BEST" != "$BESTFULL" ]; then export ADSB_PREFETCH_TILES=${BESTFULL#*@}; fi
if [ "${BEST%\%*}" != "$BEST" ]; then export ADSB_SCAN_THREADS=${BEST#*%}; BEST=${BEST%\%*}; fi
export ADSB_B200_LIB=$PWD/gr_adsb_b200/csrc/variants/libadsb_$BEST.so
el "best=$BESTFULL"
timeout 220 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_best.json 2> gpurun_out/bench_best.log
el "bench rc=$?"
timeout 260 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_best.log 2>&1
el "pytest rc=$? $(tail -1 gpurun_out/pytest_best.log)"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_slice|gather_kernel|chain_kernel|compact_kernel|init_header' -s 25 -c 40 --csv --log-file gpurun_out/launches_best.csv python bench.py --steps 12 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_list.log 2>&1
el "ncu list rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:scan_slice -s 6 -c 1 -f -o gpurun_out/scan_best python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
el "ncu full rc=$?"
