#!/bin/bash
# end-of-round bench lines (run on the GPU box): the default command, every workload, the reference arm; then the final
# k_pack4 capture
cd "$(dirname "$0")/.."
O=gpurun_out
set -x
python bench.py > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err
python bench.py --impl reference > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err
for w in c1 c3 c4 c5; do python bench.py --workload $w > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err; done
python bench.py --workload mono > $O/r02_bench_mono.json 2> $O/r02_bench_mono.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02_smoke.log 2>&1
B3="python bench.py --no-cpu --no-e2e --no-config5 --steps 2 --warmup 3 --workload c3"
ncu --clock-control none --set full --import-source on -k regex:k_pack4 -s 4 -c 1 -o $O/r02c_c3_k_pack4 $B3 > $O/r02c_ncu.log 2>&1
python tools/profile_summary.py $O/r02c_c3_k_pack4.ncu-rep k_pack4 5e6 > $O/r02c_c3_k_pack4.txt
ncu -i $O/r02c_c3_k_pack4.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02c_c3_k_pack4.source.csv.gz
rm -f $O/r02c_c3_k_pack4.ncu-rep
tail -c 400 $O/r02_bench_c2.json
