#!/bin/bash
# ncu captures of round 2 (run on the GPU box; every command has first run to completion without ncu):
#   launch lists (gpu__time_duration) of one config-2 / config-4 / config-5 step, --set full of the dominant kernels
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --no-cpu --no-e2e --no-config5 --steps 2 --warmup 3"
NCU="ncu --clock-control none"
set -x
$B > $O/r02_plain_c2.log 2>&1 && $NCU --metrics gpu__time_duration.sum -s 60 -c 400 --csv --log-file $O/r02_c2_launches.csv $B > $O/r02_ncu_l2.log 2>&1
$NCU --set full --import-source on -k regex:k_canon_s3 -s 4 -c 1 -o $O/r02_c2_s3 $B > $O/r02_ncu_a.log 2>&1
$NCU --set full -k regex:k_table_insert -s 4 -c 1 -o $O/r02_c2_insert $B > $O/r02_ncu_b.log 2>&1
$NCU --set full -k regex:k_canon_w2 -s 4 -c 1 -o $O/r02_c2_w2 $B > $O/r02_ncu_b2.log 2>&1
B5="$B --workload c5"
$B5 > $O/r02_plain_c5.log 2>&1 && $NCU --set full --import-source on -k regex:k_canon_s[23] -s 4 -c 1 -o $O/r02_c5_lane $B5 > $O/r02_ncu_c.log 2>&1
B1="$B --workload c1"
$B1 > $O/r02_plain_c1.log 2>&1 && $NCU --set full -k regex:k_canon_s[23] -s 4 -c 1 -o $O/r02_c1_lane $B1 > $O/r02_ncu_d.log 2>&1
B4="$B --workload c4"
$B4 > $O/r02_plain_c4.log 2>&1 && $NCU --metrics gpu__time_duration.sum -s 60 -c 200 --csv --log-file $O/r02_c4_launches.csv $B4 > $O/r02_ncu_l4.log 2>&1
$NCU --set full --import-source on -k regex:k_canon_seg -s 4 -c 1 -o $O/r02_c4_seg $B4 > $O/r02_ncu_e.log 2>&1
$NCU --set full -k regex:k_canon_cta -s 8 -c 2 -o $O/r02_c4_cta $B4 > $O/r02_ncu_f.log 2>&1
B3="$B --workload c3"
$B3 > $O/r02_plain_c3.log 2>&1 && $NCU --set full -k regex:k_canon_warp -s 4 -c 1 -o $O/r02_c3_warp4 $B3 > $O/r02_ncu_g.log 2>&1
$NCU --set full -k regex:k_prepare -s 4 -c 1 -o $O/r02_c3_prepare $B3 > $O/r02_ncu_h.log 2>&1
P="python tools/peer_single_rank.py"
$P > $O/r02_plain_peer.log 2>&1 && $NCU --set full -k regex:"k_owner_scatter_peers|k_table_insert_regions|k_table_first_regions|k_gather_first" -s 4 -c 4 -o $O/r02_c5_peer $P > $O/r02_ncu_i.log 2>&1
ls -la $O/r02_*.ncu-rep
