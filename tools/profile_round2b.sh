#!/bin/bash
# second batch of round-2 captures (run on the GPU box): the config-3 kernels (k_pack4, k_canon_l4), the config-4 launch list
# after the warp-cooperative duels, and the default bench lines.  Summaries only (gpurun brings back <= 64 MiB).
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --no-cpu --no-e2e --no-config5 --steps 2 --warmup 3"
NCU="ncu --clock-control none"
summ() {
    python tools/profile_summary.py $O/$1.ncu-rep "$2" $3 > $O/$1.txt 2>> $O/r02b_summ.err
    ncu -i $O/$1.ncu-rep --page source --csv 2>/dev/null | gzip > $O/$1.source.csv.gz
    ncu -i $O/$1.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/$1.raw.csv.gz
    rm -f $O/$1.ncu-rep
}
set -x
B3="$B --workload c3"
$B3 > $O/r02b_plain_c3.log 2>&1 && $NCU --metrics gpu__time_duration.sum -s 40 -c 120 --csv --log-file $O/r02b_c3_launches.csv $B3 > $O/r02b_ncu_l3.log 2>&1
$NCU --set full --import-source on -k regex:k_canon_l4 -s 4 -c 1 -o $O/r02b_c3_k_canon_l4 $B3 > $O/r02b_ncu_a.log 2>&1; summ r02b_c3_k_canon_l4 k_canon_l4 5e6
$NCU --set full --import-source on -k regex:k_pack4 -s 4 -c 1 -o $O/r02b_c3_k_pack4 $B3 > $O/r02b_ncu_b.log 2>&1; summ r02b_c3_k_pack4 k_pack4 5e6
B4="$B --workload c4"
$B4 > $O/r02b_plain_c4.log 2>&1 && $NCU --metrics gpu__time_duration.sum -s 60 -c 200 --csv --log-file $O/r02b_c4_launches.csv $B4 > $O/r02b_ncu_l4.log 2>&1
$NCU --set full -k regex:k_canon_cta -s 8 -c 2 -o $O/r02b_c4_k_canon_cta $B4 > $O/r02b_ncu_c.log 2>&1; summ r02b_c4_k_canon_cta k_canon_cta 2e3
du -sh $O
