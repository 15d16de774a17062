#!/bin/bash
# One GPU session: GPU tests, the default bench (both arms), the ncu launch list of the bench command, full ncu
# captures of the dominant kernels (raw pages + per-source-line summaries exported here; never a bench number) and
# the per-kernel bench.  usage: bash tools/round_profile.sh <tag>   -> gpurun_out/<tag>_*
set -u
R=${1:-r2}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/${R}_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/${R}_bench_full_2p31rows.json 2> $O/${R}_bench_full.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2>> $O/${R}_bench_full.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/${R}_launches.csv \
    python bench.py --rows 268435456 --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > $O/${R}_launches_bench.log 2>&1
capture() {     # capture <name> <kernel regex> <skip> <count> <command...>
    local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o $O/${R}_$name "$@" > $O/${R}_${name}_ncu.log 2>&1
    ncu -i $O/${R}_$name.ncu-rep --page raw --csv > $O/${R}_${name}_raw.csv 2>/dev/null
    ncu -i $O/${R}_$name.ncu-rep --page source --csv --print-source cuda,sass > $O/${R}_${name}_src.csv 2>/dev/null
    python tools/ncu_lines.py $O/${R}_${name}_src.csv 60 > $O/${R}_${name}_source_lines.txt
    rm -f $O/${R}_$name.ncu-rep $O/${R}_${name}_src.csv
}
capture train_tc 'linear_train_tc_kernel' 1 1 python tools/prof_train.py
capture eval_tile 'eval_tile_kernel' 1 1 python tools/prof_train.py
if [ -z "${LIGHT:-}" ]; then       # LIGHT=1: only the two dominant kernels (the others did not change)
capture misc 'bmm_kernel|decode_onehot|unpack_counts' 3 3 python tools/prof_misc.py
ONLY=cnn CNN_ROWS=262144 capture cnn 'cnn_kernel' 1 1 python tools/bench_kernels.py
fi
timeout 400 python tools/bench_kernels.py > $O/${R}_kernel_bench.jsonl 2> $O/${R}_kernel_bench.err
tail -2 $O/${R}_pytest_gpu.log
cut -c1-600 $O/${R}_bench_full_2p31rows.json
