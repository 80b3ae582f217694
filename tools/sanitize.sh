#!/bin/bash
# compute-sanitizer over the kernel tests (small shapes) and smoke(); run under gpurun. Summaries go to gpurun_out/<tag>_sanitizer_*.txt
# usage: tools/sanitize.sh <tag> [seconds per run]
set -u
tag=${1:-r02}; lim=${2:-700}
mkdir -p gpurun_out
SEL=${SANITIZE_SEL:-'dec_linear or finish_ln or vocab_argmax or decode_attention_bf16_mma or linear_tc_bf16_plain or linear_tc2_cta_pair or attention_bf16 or attention_96 or layernorm or linear_ln_emit or linear_x3 or embed_ln or greedy_kernels or beam_kernels or tag_topk or cls_attention'}
# racecheck replays every shared-memory access: by default only what changed since the last recorded run
RSEL=${SANITIZE_RACE_SEL:-$SEL}
run() {   # name, tool args..., -- command
  local name=$1; shift
  local out=gpurun_out/${tag}_sanitizer_${name}.txt
  ( timeout "$lim" compute-sanitizer "$@" ) > "$out.full" 2>&1
  local rc=$?
  { echo "# compute-sanitizer $* (exit code $rc; 124 = stopped at the ${lim} s limit)";
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error|Error|hazard|Invalid|Race" "$out.full" | sort | uniq -c | sort -rn | head -40; } > "$out"
  tail -3 "$out"
}
run memcheck_kernels --tool memcheck --print-limit 5 python -m pytest tests/test_kernels_gpu.py tests/test_decode_kernels_gpu.py -q -x -k "$SEL"
run memcheck_smoke --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()"
run racecheck_kernels --tool racecheck --racecheck-report analysis --print-limit 5 python -m pytest tests/test_kernels_gpu.py tests/test_decode_kernels_gpu.py -q -x -k "$RSEL"
run racecheck_smoke --tool racecheck --racecheck-report analysis --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()"
