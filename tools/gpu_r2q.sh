#!/bin/bash
# Round-2 pass Q: k_cbc register budget A/B (launch bounds 128 x {1, 3, 4} blocks per SM), isolated kernel times from an ncu launch list.
mkdir -p gpurun_out
L=gpurun_out/r2q.log
: > $L
for v in default cbc3 cbc4; do
  if [ $v = default ]; then unset MFC5_LIB; else export MFC5_LIB=$PWD/mfc_b200/libmfc5_$v.so; fi
  echo "== $v" | tee -a $L
  timeout 300 python tools/sweep_bench.py --steps 10 2>&1 | tail -1 | tee -a $L
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_cbc -s 20 -c 12 --csv --log-file gpurun_out/cbc_$v.csv python tools/sweep_bench.py --steps 3 > /dev/null 2>&1
  python - <<PY | tee -a $L
import csv
rows=[r for r in csv.reader(open("gpurun_out/cbc_$v.csv")) if r and r[0].isdigit()]
d=[float(r[-1].replace(",","")) for r in rows]
u=rows[0][-2] if rows else ""
print("k_cbc launches", len(d), "durations", sorted(set(round(x,1) for x in d))[:6], u)
PY
done
