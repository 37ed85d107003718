#!/bin/bash
# compute-sanitizer over the smoke path and the co-scheduled kernel test.  NOTE: memcheck under the python + torch start-up is
# very slow (no output after 15 min on the B200 box); every step is bounded to 5 min so that the call cannot eat the budget.
OUT=gpurun_out
mkdir -p $OUT
S=/usr/local/cuda/bin/compute-sanitizer
( echo "== memcheck: __graft_entry__.smoke()"; timeout 300 $S --tool memcheck --print-limit 5 python __graft_entry__.py smoke 2>&1 | grep -E "smoke ok|ERROR SUMMARY|Invalid|Error" | head -8
  echo "== racecheck: __graft_entry__.smoke()"; timeout 300 $S --tool racecheck --print-limit 5 python __graft_entry__.py smoke 2>&1 | grep -E "smoke ok|RACECHECK SUMMARY|hazard|Error" | head -8
  echo "== memcheck: tests -k 'coscheduled and case0'"; timeout 300 $S --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "coscheduled and case0" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | head -8
  echo "== racecheck: tests -k 'coscheduled and case0'"; timeout 300 $S --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "coscheduled and case0" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | head -8
  echo "== memcheck: fusion V=257,400 (multi-lane kernels)"; timeout 300 $S --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fusion_bit_exact and (257 or 400)" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | head -8
) | tee $OUT/r2b_sanitizer.txt
