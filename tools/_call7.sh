set -u
mkdir -p gpurun_out
exec 2>&1
T0=$(date +%s)
el() { echo "[+$(( $(date +%s) - T0 ))s] $*"; }
timeout 200 ncu --set full --clock-control none --import-source on -k regex:scan_slice -s 6 -c 1 -f -o gpurun_out/scan_r2g python bench.py --steps 4 --warmup 3 --quick --no-verify > gpurun_out/ncu_full_r2g.log 2>&1
el "ncu full rc=$?"
tail -3 gpurun_out/ncu_full_r2g.log
