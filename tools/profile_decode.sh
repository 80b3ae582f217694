#!/bin/bash
# ncu --set full captures of single decode-step kernels (run under gpurun). Launch order inside one eager decode step, counting
# only the kernels the regex matches: embed 0 | per layer: qkv 1, attention 2, o-proj 3, finish 4, fc1 5, fc2 6, finish 7 |
# head transform 29, finish 30, vocabulary arg-max 31, token step 32.
# usage: tools/profile_decode.sh <tag> <name:index> ...      e.g.  tools/profile_decode.sh r02 fc1:5 vocab:31
set -u
tag=$1; shift
mkdir -p gpurun_out
RX='regex:gemm_dec|finish_ln|token_step|decode_att|embed_ln'
for spec in "$@"; do
  name=${spec%%:*}; idx=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k "$RX" --launch-skip "$idx" -c 1 -f \
      -o "gpurun_out/${tag}_dec_${name}" python tools/decode_probe.py 512 eager > "gpurun_out/${tag}_dec_${name}.log" 2>&1
  echo "== $name (launch $idx): rc=$?"
done
