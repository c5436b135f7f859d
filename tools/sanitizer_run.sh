#!/bin/bash
# compute-sanitizer over a small slice of the gpu suite (memcheck: every kernel family; racecheck: the latency path, the
# peer-mode step and the scan kernel's shared-memory phases).  Slow (10-50x): minutes on a B200.
#   bash tools/sanitizer_run.sh > profiles/r02/compute_sanitizer.txt
cd "$(dirname "$0")/.."
SEL='test_peer_shards_equal_whole_buffer and (2-2-1 or 8-2-4) or test_peer_carry_cascade or test_latency_path_equals_throughput_pipeline and (2-f32-8192 or 8-c64-65536) or test_latency_path_falls_back or test_stream_equals_whole_buffer and 2-1-200000-16384-8192 or test_stream_integer_formats or test_demod_uses_the_framers_bits'
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Race reported|hazard" | head -20
  echo "exit=$?"
done
