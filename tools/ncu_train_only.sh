set -u
O=gpurun_out; R=r2e
capture() { local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o $O/${R}_$name "$@" > $O/${R}_${name}_ncu.log 2>&1
  ncu -i $O/${R}_$name.ncu-rep --page raw --csv > $O/${R}_${name}_raw.csv 2>/dev/null
  ncu -i $O/${R}_$name.ncu-rep --page source --csv --print-source cuda,sass > $O/${R}_${name}_src.csv 2>/dev/null
  python tools/ncu_lines.py $O/${R}_${name}_src.csv 80 > $O/${R}_${name}_source_lines.txt
  rm -f $O/${R}_$name.ncu-rep $O/${R}_${name}_src.csv; }
capture train_tc 'linear_train_tc_kernel' 1 1 python tools/prof_train.py
