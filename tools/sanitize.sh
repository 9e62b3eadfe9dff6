#!/bin/bash
# compute-sanitizer over tools/sanitize_smoke.py (GPU box): memcheck on the full-size smoke, racecheck + initcheck on a
# reduced one. Logs -> gpurun_out/sanitize_*.log; the summaries are what profiles/sanitizer_*.md quotes.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
T=${SAN_TIMEOUT:-420}
timeout $T $SAN --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit $?" | tee -a gpurun_out/sanitize_memcheck.log
RTDS_SAN_SCALE=0.25 timeout $T $SAN --tool racecheck --racecheck-report all --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit $?" | tee -a gpurun_out/sanitize_racecheck.log
RTDS_SAN_SCALE=0.25 timeout $T $SAN --tool initcheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/sanitize_initcheck.log 2>&1
echo "initcheck exit $?" | tee -a gpurun_out/sanitize_initcheck.log
tail -n 5 gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_racecheck.log gpurun_out/sanitize_initcheck.log
