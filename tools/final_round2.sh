#!/bin/bash
# end-of-round bench lines (run on the GPU box): the default command, every workload, the reference arm; then the final
# k_pack4 capture
cd "$(dirname "$0")/.."
O=gpurun_out
set -x
python -m pytest tests -m gpu -q > $O/r02_pytest_gpu.log 2>&1
python bench.py > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err
python bench.py --impl reference > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err
for w in c1 c3 c4 c5; do python bench.py --workload $w > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err; done
python bench.py --workload mono > $O/r02_bench_mono.json 2> $O/r02_bench_mono.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02_smoke.log 2>&1
B3="python bench.py --no-cpu --no-e2e --no-config5 --steps 2 --warmup 3 --workload c3"
for k in k_pack4 k_canon_l4; do
ncu --clock-control none --set full --import-source on -k regex:$k -s 4 -c 1 -o $O/r02d_c3_$k $B3 > $O/r02d_ncu_$k.log 2>&1
python tools/profile_summary.py $O/r02d_c3_$k.ncu-rep $k 5e6 > $O/r02d_c3_$k.txt
ncu -i $O/r02d_c3_$k.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02d_c3_$k.source.csv.gz
rm -f $O/r02d_c3_$k.ncu-rep
done
ncu --clock-control none --metrics gpu__time_duration.sum -s 40 -c 120 --csv --log-file $O/r02d_c3_launches.csv $B3 > $O/r02d_ncu_l3.log 2>&1
tail -c 400 $O/r02_bench_c2.json
