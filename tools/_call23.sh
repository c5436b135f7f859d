exec 2>&1
mkdir -p gpurun_out
bash tools/sanitizer_run.sh > gpurun_out/compute_sanitizer.txt 2>&1
cat gpurun_out/compute_sanitizer.txt | cut -c1-250
