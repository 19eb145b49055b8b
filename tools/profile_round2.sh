#!/bin/bash
# ncu captures of round 2 (run on the GPU box; every command has first run to completion without ncu):
#   launch lists (gpu__time_duration) of one config-2 / config-4 step, --set full of the dominant kernels.
# gpurun brings back at most 64 MiB, so every report is summarised here (tools/profile_summary.py, raw / source CSV gzipped) and
# only the hot kernel's .ncu-rep is kept.
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --no-cpu --no-e2e --no-config5 --steps 2 --warmup 3"
NCU="ncu --clock-control none"
summ() {   # summ <rep-stem> <kernel pattern> <units per launch> [keep]
    python tools/profile_summary.py $O/$1.ncu-rep "$2" $3 > $O/$1.txt 2>> $O/r02_summ.err
    ncu -i $O/$1.ncu-rep --page source --csv 2>/dev/null | gzip > $O/$1.source.csv.gz
    ncu -i $O/$1.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/$1.raw.csv.gz
    [ -z "$4" ] && rm -f $O/$1.ncu-rep
}
set -x
$B > $O/r02_plain_c2.log 2>&1 && $NCU --metrics gpu__time_duration.sum -s 60 -c 400 --csv --log-file $O/r02_c2_launches.csv $B > $O/r02_ncu_l2.log 2>&1
$NCU --set full --import-source on -k regex:k_canon_s3 -s 4 -c 1 -o $O/r02_c2_k_canon_s3 $B > $O/r02_ncu_a.log 2>&1; summ r02_c2_k_canon_s3 k_canon_s3 1e7 keep
$NCU --set full -k regex:k_table_insert -s 4 -c 1 -o $O/r02_c2_k_table_insert $B > $O/r02_ncu_b.log 2>&1; summ r02_c2_k_table_insert k_table_insert 1e7
$NCU --set full -k regex:k_canon_w2 -s 4 -c 1 -o $O/r02_c2_k_canon_w2 $B > $O/r02_ncu_b2.log 2>&1; summ r02_c2_k_canon_w2 k_canon_w2 1e7
B5="$B --workload c5"
$B5 > $O/r02_plain_c5.log 2>&1 && $NCU --set full --import-source on -k 'regex:k_canon_s[23]' -s 4 -c 1 -o $O/r02_c5_lane $B5 > $O/r02_ncu_c.log 2>&1; summ r02_c5_lane k_canon_s 1.25e7
B1="$B --workload c1"
$B1 > $O/r02_plain_c1.log 2>&1 && $NCU --set full -k 'regex:k_canon_s[23]' -s 4 -c 1 -o $O/r02_c1_lane $B1 > $O/r02_ncu_d.log 2>&1; summ r02_c1_lane k_canon_s 1e6
B4="$B --workload c4"
$B4 > $O/r02_plain_c4.log 2>&1 && $NCU --metrics gpu__time_duration.sum -s 60 -c 200 --csv --log-file $O/r02_c4_launches.csv $B4 > $O/r02_ncu_l4.log 2>&1
$NCU --set full --import-source on -k regex:k_canon_seg -s 4 -c 1 -o $O/r02_c4_k_canon_seg $B4 > $O/r02_ncu_e.log 2>&1; summ r02_c4_k_canon_seg k_canon_seg 2e5
$NCU --set full -k regex:k_canon_cta -s 8 -c 2 -o $O/r02_c4_k_canon_cta $B4 > $O/r02_ncu_f.log 2>&1; summ r02_c4_k_canon_cta k_canon_cta 2e3
B3="$B --workload c3"
$B3 > $O/r02_plain_c3.log 2>&1 && $NCU --set full -k regex:k_canon_warp -s 4 -c 1 -o $O/r02_c3_k_canon_warp4 $B3 > $O/r02_ncu_g.log 2>&1; summ r02_c3_k_canon_warp4 k_canon_warp 5e6
$NCU --set full -k regex:k_prepare -s 4 -c 1 -o $O/r02_c3_k_prepare $B3 > $O/r02_ncu_h.log 2>&1; summ r02_c3_k_prepare k_prepare 5e6
P="python tools/peer_single_rank.py"
$P > $O/r02_plain_peer.log 2>&1 && $NCU --set full -k 'regex:k_owner_scatter_peers|k_table_insert_regions|k_table_first_regions|k_gather_first' -s 4 -c 4 -o $O/r02_c5_peer_kernels $P > $O/r02_ncu_i.log 2>&1; summ r02_c5_peer_kernels k_ 1.25e7
du -sh $O; ls -la $O | head -60
