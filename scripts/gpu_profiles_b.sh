#!/bin/bash
# Late round-2 ncu evidence (run under gpurun, ONE call): launch lists of the NW-UCLA and ST-GCN training steps and
# `--set full` captures of the kernels added late in the round.  Every ncu command is preceded by the same command run
# plainly; ncu only runs if that exits 0.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
python scripts/profile_step.py bf16 64 > $O/pb_step_plain.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum --profile-from-start off --cache-control none --csv --log-file $O/r02b_step_launches_warm.csv python scripts/profile_step.py bf16 64 > $O/pb_step_ncu.log 2>&1
python scripts/profile_step.py bf16 16 stgcn_train > $O/pb_ststep_plain.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum --profile-from-start off --cache-control none --csv --log-file $O/r02b_stgcn_step_launches_warm.csv python scripts/profile_step.py bf16 16 stgcn_train > $O/pb_ststep_ncu.log 2>&1
python scripts/profile_conv.py 32 64 64 300 25 9 > $O/pb_t9_plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:tconv9 -s 9 -c 3 -f -o $O/r02b_tconv9 python scripts/profile_conv.py 32 64 64 300 25 9 > $O/pb_t9_ncu.log 2>&1
python scripts/profile_ctrgc.py bf16 ntu_l8 > $O/pb_l8_plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:ctrgc -s 6 -c 2 -f -o $O/r02b_ctrgc_ntu_l8 python scripts/profile_ctrgc.py bf16 ntu_l8 > $O/pb_l8_ncu.log 2>&1
python scripts/profile_ctrgc.py bf16 ntu2048 > $O/pb_n2048_plain.log 2>&1 &&
$NCU --set full -k regex:ctrgc_fwd -s 3 -c 1 -f -o $O/r02b_ctrgc_ntu2048 python scripts/profile_ctrgc.py bf16 ntu2048 > $O/pb_n2048_ncu.log 2>&1
python scripts/bench_stgcn_kernels.py agg > $O/pb_agg_plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:graph_agg -s 9 -c 3 -f -o $O/r02b_graph_agg python scripts/bench_stgcn_kernels.py agg > $O/pb_agg_ncu.log 2>&1
ls -la $O/*.ncu-rep | tail
