"""Known-answer vectors of `monomerize`, taken from the reference's own unit tests (lib/src/monomerize.rs:155-554).

Run in the build container (needs /root/reference for the two long literals of the `ambivirus` test, which are read
from the reference file rather than retyped); writes tests/golden/monomerize_kats.json.  Every vector names the test and
line it comes from.  Expected values are the reference's asserted outputs, not the oracle's.
"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/lib/src/monomerize.rs"
src = open(REF).read()

kats = []


def add(name, line, seq, expected, seed_len, overlap_dist=None, min_identity=None, sensitive=False):
    kats.append(dict(name=name, ref="lib/src/monomerize.rs:%d" % line, seq=seq, expected=expected, seed_len=seed_len,
                     overlap_dist=overlap_dist, min_identity=min_identity, sensitive=sensitive))


# basic_tests::monomerize_with_seed_len_and_overlap_dist_works (:176-199)
for seq, exp, s, d in (("ATGCATGC", "ATGC", 4, 0), ("ATGCATGC", "ATGC", 4, 1), ("ATGCATGC", "ATGC", 2, 0), ("ATGCATGC", "ATGC", 2, 1),
                       ("AAAAATTTTTAAAAATTTTT", "AAAAATTTTT", 10, 0)):
    add("seed_len_and_overlap_dist_works", 176, seq, exp, s, d)
# no_overlap (:201-214)
for s in range(4, 11):
    for d in (0, 1, 2):
        add("no_overlap", 201, "TTTTTTTTTTTTAAAAAAAAAA", "TTTTTTTTTTTTAAAAAAAAAA", s, d)
# mismatches_within_limit (:216-231)
x = "TTAGCCCGTGTTTTATCGGAAGCTATCCTCAAAGCCCGTGTTTTATCGGAAGCTATCCTC"
for s in range(4, 11):
    for d in (2, 3, 4):
        add("mismatches_within_limit", 216, x, "TTAGCCCGTGTTTTATCGGAAGCTATCCTC", s, d)
# too_many_mismatches (:233-248)
x = "TTTTTGGTTTTTAAAAAAAAAATTTTTTTTTTTTAAAAAAAAAA"
for s in range(4, 11):
    for d in (0, 1):
        add("too_many_mismatches", 233, x, x, s, d)
# complete_identical_multimer (:250-264)
for s in range(4, 11):
    for d in (0, 1, 2):
        add("complete_identical_multimer", 250, "AAAAATTTTTAAAAATTTTTAAAAATTTTT", "AAAAATTTTT", s, d)
# multimer_with_mismatch_in_each (:266-283)
for s in (4, 5, 6, 7):
    for d in (1, 2, 3):
        add("multimer_with_mismatch_in_each", 266, "AACAATTTTTAAGAATTTTTAAAAATTTTT", "AACAATTTTT", s, d)
# multimer_with_mismatch_in_middle (:285-298)
add("multimer_with_mismatch_in_middle", 285, "AAAAATTTTTAAGAATTTTTAAAAATTTTT", "AAAAATTTTT", 5, 1)
# multimer_with_seed_repeated (:315-328)
add("multimer_with_seed_repeated", 315, "TGCCAATGCATGCCAATGC", "TGCCAATGCA", 4, 0)
# big_dimer (:330-349)
for s in range(6, 12):
    for d in (0, 1, 3):
        add("big_dimer", 330, "ATGACAGGTACAGCATAATGACAGGTACAGCATA", "ATGACAGGTACAGCATA", s, d)
# ambivirus (:351-372): the two literals are read from the reference file
m = re.search(r'fn ambivirus\(.*?let input = b"([ACGT]+)";\s*let monomer = b"([ACGT]+)";', src, re.S)
amb_in, amb_mono = m.group(1), m.group(2)
for s in (10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 32, 33, 63):
    for d in range(0, 11):
        add("ambivirus", 351, amb_in, amb_mono, s, d)
# single_pass_regressions (:374-391): min identity 0.95, seed 10
for seq, exp in (
    ("TCCTCCATCACCTAGTTTATGTAGAAACGCTGCTAAATCAATTTCCTCCATCACCTAGTTTATGTAGAAAAGCTGCTAAATCAATTTCCTCCATCACCTAGTTTATGTAGAAACGCTGCTAAATCAATTTCCTCCATCACCTAGTTTATGTAGAAAAGCTGCTA",
     "TCCTCCATCACCTAGTTTATGTAGAAACGCTGCTAAATCAATT"),
    ("GCAGTTATAGAGAGAGTGGGTCAGTTCATTATTACACTGCAGTAATAGAGAGAGTGGGTCAGTTCATTATTACACTGCAGTGATAGAGAGAGTGGGTCAGTTCATTATTACACTGCAGTTATAGAGAGAGTGGGTCAGTTCATTATTACACTGCAGTAATAGAGAGAGTGGGTCAGTTCATTATTACACTGCAGTGATAG",
     "GCAGTTATAGAGAGAGTGGGTCAGTTCATTATTACACT"),
    ("CTGGCCCAGGGGCTTCTAGTCAAACAGGCCTCTCTTCCCCACTCCTTACCTCTTCTGGTCTCTGGCCCAGGGGCTTCTAGTCAAACAGGCCTCTCTTCCCCACTCCTTACCTCTTCTGGTCTCTGGCCCCTGGCCCAGGGGCTTCTAGTCAAACAGGCCTCTCTTCCCCACTCCTTACCTCTTCTGGTCTCTGGCCCAGGGGCTTCT",
     "CTGGCCCAGGGGCTTCTAGTCAAACAGGCCTCTCTTCCCCACTCCTTACCTCTTCTGGTCT")):
    add("single_pass_regressions", 374, seq, exp, 10, None, 0.95)
# overlap_percent::dimer_with_overlap_percentage (:424-446)
x = "ATGCCCATGCGCCAGCGCAGATGCGAATGCGCCAGCGCAG"
add("dimer_with_overlap_percentage", 424, x, x, 4, None, 0.95)
add("dimer_with_overlap_percentage", 424, x, "ATGCCCATGCGCCAGCGCAG", 4, None, 0.90)
# overlap_percentage_rounds_down_to_nearest_nt (:448-477)
x = "TGCCCATGCGCCAGCGCAGATGCGAATGCGCCAGCGCAG"
add("overlap_percentage_rounds_down", 448, x, x, 4, None, 0.95)
add("overlap_percentage_rounds_down", 448, x, "TGCCCATGCGCCAGCGCAGA", 4, None, 0.90)
add("overlap_percentage_rounds_down", 448, x, "TGCCCATGCGCCAGCGCAGA", 4, None, 0.94)
# sensitive::sensitive_monomerization (:499-510)
add("sensitive_monomerization", 499, "ATGCCCATGCGCCAGCGCAAATGCCCATGCGCCAGCGCAG", "ATGCCCATGCGCCAGCGCAA", 4, None, 0.95, True)

strings = sorted({k["seq"] for k in kats} | {k["expected"] for k in kats})      # every literal once
index = {x: i for i, x in enumerate(strings)}
for k in kats:
    k["seq"], k["expected"] = index[k["seq"]], index[k["expected"]]
json.dump({"strings": strings, "kats": kats}, open(os.path.join(HERE, "monomerize_kats.json"), "w"), indent=0)
print(len(kats), "vectors,", len(strings), "distinct strings")
