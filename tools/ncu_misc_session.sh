#!/bin/bash
# ncu --set full capture of the streaming kernels (bmm, decode_onehot, unpack_counts); raw page only.
set -u
O=gpurun_out
T=${TAG:-misc}
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bmm_kernel|decode_onehot|unpack_counts' -s 3 -c 3 \
    -o $O/${T} python tools/prof_misc.py > $O/${T}_ncu.log 2>&1
ncu -i $O/${T}.ncu-rep --page raw --csv > $O/${T}_raw.csv 2>/dev/null
ncu -i $O/${T}.ncu-rep --page source --csv --print-source cuda,sass > $O/${T}_src.csv 2>/dev/null
python tools/ncu_lines.py $O/${T}_src.csv 25 > $O/${T}_lines.txt
rm -f $O/${T}.ncu-rep $O/${T}_src.csv
tail -2 $O/${T}_ncu.log
