#!/bin/bash
# One GPU session: GPU tests, the default bench (both arms), the ncu launch list of the bench command, full ncu
# captures of the dominant kernels (raw pages exported here; never a bench number) and the per-kernel bench.
set -u
R=${1:-r1}
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/${R}_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/${R}_bench_full.json 2> $O/${R}_bench_full.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2>> $O/${R}_bench_full.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${R}_launches.csv \
    python bench.py --rows 268435456 --steps 2 --warmup 3 --no-cpu-baseline > $O/${R}_launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"linear_train2_kernel|eval_kernel" -s 2 -c 2 \
    -o $O/${R}_fused python tools/prof_train.py > $O/${R}_ncu_fused.log 2>&1
ncu -i $O/${R}_fused.ncu-rep --page raw --csv > $O/${R}_fused_raw.csv 2>/dev/null
ncu -i $O/${R}_fused.ncu-rep --page source --csv --print-source cuda,sass > $O/${R}_fused_src.csv 2>/dev/null
rm -f $O/${R}_fused.ncu-rep
ONLY=cnn CNN_ROWS=262144 timeout 300 ncu --set full --clock-control none --import-source on -k regex:cnn_kernel -s 1 -c 1 \
    -o $O/${R}_cnn python tools/bench_kernels.py > $O/${R}_ncu_cnn.log 2>&1
ncu -i $O/${R}_cnn.ncu-rep --page raw --csv > $O/${R}_cnn_raw.csv 2>/dev/null
rm -f $O/${R}_cnn.ncu-rep
timeout 300 python tools/bench_kernels.py > $O/${R}_kernel_bench.jsonl 2> $O/${R}_kernel_bench.err
tail -2 $O/${R}_pytest.log
cat $O/${R}_bench_full.json | cut -c1-400
