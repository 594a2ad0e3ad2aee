#!/usr/bin/env bash
# compute-sanitizer targets: memcheck by default, `bench/sanitize.sh racecheck|synccheck|initcheck` for the rest.
# One tiny launch per kernel family (bench/sanitize_smoke.py); the report lands in gpurun_out/sanitizer_<tool>.log.
set -u
tool="${1:-memcheck}"
mkdir -p gpurun_out
compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 python bench/sanitize_smoke.py > "gpurun_out/sanitizer_${tool}.log" 2>&1
rc=$?
tail -5 "gpurun_out/sanitizer_${tool}.log"
exit $rc
