#!/bin/bash
# ncu --set full capture of one train-kernel launch (KERNEL, default linear_train_tc_kernel) on a synthetic table of regime $REGIME; exports the raw
# page and the per-source-line aggregation.
set -u
O=gpurun_out
T=${TAG:-c}
mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-linear_train_tc_kernel} -s 1 -c 1 \
    -o $O/${T} python tools/prof_train.py > $O/${T}_ncu.log 2>&1
ncu -i $O/${T}.ncu-rep --page raw --csv > $O/${T}_raw.csv 2>/dev/null
ncu -i $O/${T}.ncu-rep --page source --csv --print-source cuda,sass > $O/${T}_src.csv 2>/dev/null
python tools/ncu_lines.py $O/${T}_src.csv 60 > $O/${T}_lines.txt
rm -f $O/${T}.ncu-rep $O/${T}_src.csv
tail -3 $O/${T}_ncu.log
head -30 $O/${T}_lines.txt
