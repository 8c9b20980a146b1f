#!/usr/bin/env bash
# compute-sanitizer over the kernels of the hot path.  Run on the GPU box:
#   bash tools/sanitize.sh [outdir]          (1 GPU: memcheck, racecheck, synccheck over tools/sanitize_workload.py)
#   bash tools/sanitize.sh [outdir] peer     (2 GPUs: memcheck + racecheck of the peer exchanges, one sanitizer per rank;
#                                             --report-api-errors no: torch's symmetric-memory set-up PROBES for fabric handles with
#                                             a cuMemCreate that is allowed to fail, which memcheck would count as an error)
# Only this library's kernels are instrumented (--kernel-name kns=dsnt); logs go to <outdir>/sanitizer_<tool>_<case>.log and a
# one-line verdict per run to <outdir>/sanitizer_summary.txt.
set -uo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
OUT="${1:-$ROOT/gpurun_out}"
MODE="${2:-single}"
mkdir -p "$OUT"
SUM="$OUT/sanitizer_summary.txt"
cd "$ROOT"
run_one() {   # tool, label, command...
  local tool="$1" label="$2"; shift 2
  local log="$OUT/sanitizer_${tool}_${label}.log"
  timeout 1500 compute-sanitizer --tool "$tool" --kernel-name kns=dsnt --error-exitcode 9 --launch-timeout 120 \
      --print-limit 20 "$@" > "$log" 2>&1
  local rc=$?
  local errs
  errs="$(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$log" | tail -1)"
  echo "$tool $label rc=$rc ${errs:-no summary line} done=$(grep -c SANITIZE-WORKLOAD-DONE "$log")" | tee -a "$SUM"
}
if [ "$MODE" = "peer" ]; then
  for tool in memcheck racecheck; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
      --no-python compute-sanitizer --tool "$tool" --kernel-name kns=dsnt --error-exitcode 9 --launch-timeout 120 \
      --report-api-errors no --print-limit 20 --log-file "$OUT/sanitizer_${tool}_peer_%q{RANK}.log" python tools/sanitize_workload.py peer \
      > "$OUT/sanitizer_${tool}_peer_stdout.log" 2>&1
    rc=$?
    for r in 0 1; do
      echo "$tool peer rank $r rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/sanitizer_${tool}_peer_${r}.log" | tail -1)" | tee -a "$SUM"
    done
  done
else
  # one process per tool (importing torch under the sanitizer is the slow part): memcheck over every kernel family,
  # racecheck / synccheck over the kernels with shared-memory protocols (mbarrier rings, DSMEM exchange, block reductions)
  run_one memcheck all python tools/sanitize_workload.py step stacked generic pair two level1
  run_one racecheck smem python tools/sanitize_workload.py step stacked generic pair two
  run_one synccheck smem python tools/sanitize_workload.py step stacked generic pair two
fi
