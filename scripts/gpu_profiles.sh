#!/bin/bash
# Round-2 ncu evidence (run under gpurun, ONE call): launch lists of a training step (cold / warm caches) and
# `--set full` captures of the dominant kernels.  Every ncu command is preceded by the same command run plainly.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
python scripts/profile_step.py bf16 64 > $O/p_step_plain.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file $O/r02_step_launches.csv python scripts/profile_step.py bf16 64 > $O/p_step_ncu.log 2>&1
$NCU --metrics gpu__time_duration.sum --profile-from-start off --cache-control none --csv --log-file $O/r02_step_launches_warm.csv python scripts/profile_step.py bf16 64 > $O/p_step_ncu2.log 2>&1
python scripts/profile_ctrgc.py bf16 ucla > $O/p_ctrgc_ucla_plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:ctrgc -s 6 -c 2 -f -o $O/r02_ctrgc_ucla python scripts/profile_ctrgc.py bf16 ucla > $O/p_ctrgc_ucla_ncu.log 2>&1
python scripts/profile_ctrgc.py bf16 ntu > $O/p_ctrgc_ntu_plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:ctrgc -s 6 -c 2 -f -o $O/r02_ctrgc_ntu python scripts/profile_ctrgc.py bf16 ntu > $O/p_ctrgc_ntu_ncu.log 2>&1
python scripts/profile_conv.py 64 64 192 52 20 1 > $O/p_conv_plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:conv_ -s 10 -c 3 -f -o $O/r02_conv_l2 python scripts/profile_conv.py 64 64 192 52 20 1 > $O/p_conv_ncu.log 2>&1
python scripts/profile_conv.py 64 16 16 52 20 5 > $O/p_tconv_plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:conv -s 9 -c 3 -f -o $O/r02_tconv_l2 python scripts/profile_conv.py 64 16 16 52 20 5 > $O/p_tconv_ncu.log 2>&1
python scripts/profile_conv.py 32 64 64 300 25 9 > $O/p_stconv_plain.log 2>&1 &&
$NCU --set full --import-source on -k regex:conv -s 9 -c 3 -f -o $O/r02_stgcn_conv9 python scripts/profile_conv.py 32 64 64 300 25 9 > $O/p_stconv_ncu.log 2>&1
ls -la $O/*.ncu-rep | tail
