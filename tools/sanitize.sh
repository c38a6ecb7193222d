#!/bin/bash
# compute-sanitizer over small cases of every kernel family (tools/sanitize_cases.py): memcheck, racecheck (shared-memory
# hazards of the warp-specialised mbarrier pipelines and the cp.async residual ring) and synccheck (barrier misuse).
# Usage (on a B200 box):  bash tools/sanitize.sh [outdir]      -> <outdir>/sanitize_{memcheck,racecheck,synccheck}.log
# Each tool runs under its own timeout: the sanitizer serialises kernels and instruments every shared-memory access.
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
status=0
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool" | tee "$OUT/sanitize_$tool.log"
  timeout ${SANITIZE_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
      python tools/sanitize_cases.py >> "$OUT/sanitize_$tool.log" 2>&1
  rc=$?
  echo "== exit code $rc" | tee -a "$OUT/sanitize_$tool.log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|exact|finite" "$OUT/sanitize_$tool.log" | tail -20
  [ $rc -ne 0 ] && status=$rc
done
exit $status
