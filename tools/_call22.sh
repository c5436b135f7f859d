exec 2>&1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()"
echo "smoke rc=$?"
bash tools/sanitizer_run.sh > gpurun_out/compute_sanitizer.txt 2>&1
cat gpurun_out/compute_sanitizer.txt
