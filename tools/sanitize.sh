#!/bin/sh
# compute-sanitizer over the hot path: memcheck + racecheck + synccheck of the smoke test (fp32 and
# tcgen05 modes, SSG) and of one MSG attack step -- the fused kernels alias scratch over live operand buffers by design, and
# the round-1 bug was a shared-memory race.  Run on a GPU box from the repo root:
#     sh tools/sanitize.sh            # logs under gpurun_out/sanitize_*.log, summary to stdout
# Sizes are small (the tools slow kernels 10-100x); racecheck sees shared memory only, memcheck everything.
set -u
OUT="${1:-gpurun_out}"
mkdir -p "$OUT"
PY="${PYTHON:-python}"
status=0
for tool in memcheck racecheck synccheck; do
  log="$OUT/sanitize_$tool.log"
  timeout 700 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
      "$PY" tools/sanitize_workload.py > "$log" 2>&1
  rc=$?
  tail -n 3 "$log" | sed "s/^/[$tool] /"
  echo "[$tool] exit code $rc"
  [ "$rc" -eq 0 ] || status=1
done
exit $status
