#!/bin/bash
# compute-sanitizer over the GPU parity tests (run under gpurun): memcheck, racecheck, synccheck -> gpurun_out/<tag>_sanitizer_*.log
# usage: tools/sanitize.sh <tag>
cd "$(dirname "$0")/.."
TAG=${1:-san}
O=gpurun_out
mkdir -p $O
(compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mask_break.py tests/test_gpu_ingest.py tests/test_gpu_emit.py tests/test_evaluator.py -q -x -m gpu > $O/${TAG}_sanitizer_memcheck.log 2>&1; echo memcheck rc=$? >> $O/${TAG}_sanitizer_memcheck.log); tail -4 $O/${TAG}_sanitizer_memcheck.log
(compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mask_break.py -q -x -m gpu -k "adversarial or long_reads or golden" > $O/${TAG}_sanitizer_racecheck.log 2>&1; echo racecheck rc=$? >> $O/${TAG}_sanitizer_racecheck.log); tail -4 $O/${TAG}_sanitizer_racecheck.log
(compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "adversarial and 3" > $O/${TAG}_sanitizer_synccheck.log 2>&1; echo synccheck rc=$? >> $O/${TAG}_sanitizer_synccheck.log); tail -4 $O/${TAG}_sanitizer_synccheck.log
