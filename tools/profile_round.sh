#!/bin/bash
# Profile capture of the bench command on the GPU box (run through gpurun).  Usage: tools/profile_round.sh <tag>
#   1. the plain command must exit 0 first (a number printed under ncu is never a bench value)
#   2. launch list:  ncu --metrics gpu__time_duration.sum --clock-control none   (B200_PROFILING.md)
#   3. one `--set full` capture of the dominant kernels of one transformer block (4 GEMM shapes) and of the attention kernel
set -u
TAG=${1:-r01}
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --legs \"\""
$CMD > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
# steady state: skip the first step's GEMMs (49 per step), take the four GEMMs of one block
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 102 -c 4 -f -o $OUT/${TAG}_gemm_full $CMD > $OUT/${TAG}_ncu_gemm.log 2>&1
echo "gemm full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_row -s 26 -c 1 -f -o $OUT/${TAG}_attn_full $CMD > $OUT/${TAG}_ncu_attn.log 2>&1
echo "attention full rc=$?"
ls -la $OUT | grep ${TAG}
