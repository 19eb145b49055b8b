#!/bin/bash
# after the in-lane tie resolution: the default bench line, the lane kernel's ncu summary and the launch list of a config-2 step
cd "$(dirname "$0")/.."
O=gpurun_out
set -x
python bench.py > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err
B="python bench.py --no-cpu --no-e2e --no-config5 --steps 2 --warmup 3"
$B > $O/r02e_plain_c2.log 2>&1 && ncu --clock-control none --metrics gpu__time_duration.sum -s 60 -c 400 --csv --log-file $O/r02e_c2_launches.csv $B > $O/r02e_ncu_l2.log 2>&1
ncu --clock-control none --set full --import-source on -k regex:k_canon_s3 -s 4 -c 1 -o $O/r02e_c2_k_canon_s3 $B > $O/r02e_ncu_a.log 2>&1
python tools/profile_summary.py $O/r02e_c2_k_canon_s3.ncu-rep k_canon_s3 1e7 > $O/r02e_c2_k_canon_s3.txt
ncu -i $O/r02e_c2_k_canon_s3.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02e_c2_k_canon_s3.source.csv.gz
ncu -i $O/r02e_c2_k_canon_s3.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/r02e_c2_k_canon_s3.raw.csv.gz
rm -f $O/r02e_c2_k_canon_s3.ncu-rep $O/r02_c2_k_canon_s3.ncu-rep
tail -c 300 $O/r02_bench_c2.json
