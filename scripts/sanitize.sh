#!/bin/bash
# compute-sanitizer passes over a small slice of the GPU test-suite (run under gpurun)
set -x
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "reference_routine or ragged or empty_and_tiny or score_modes_messages_sequence or split_scan" 2>&1 | tail -15
echo "memcheck rc=$?"
timeout 900 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "reference_routine_on_captures and 457780" 2>&1 | tail -15
echo "racecheck rc=$?"
timeout 600 $SAN --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "reference_routine_on_captures and 457780" 2>&1 | tail -8
echo "synccheck rc=$?"
