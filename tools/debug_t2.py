"""Debug helper: lane-per-record kernel vs oracle on a c1-like batch; prints mismatch statistics."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import circkit_b200, oracle
from circkit_b200 import device as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 250
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 400
mask = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ctx = circkit_b200.Context(max_batch_bytes=0, max_batch_records=0)
b = D.synth_batch(ctx, seed=11, first_index=0, n_records=n, kind=int(sys.argv[5]) if len(sys.argv) > 5 else 0, lo=lo, hi=hi, dup_permille=0, adversarial_permille=0)
outs = D.CanonOutputs(n, b.total, b.offsets.device, want_bytes=True, want_hash=True, aligned=True)
ws = D.Workspace(ctx, n)
D.canon_packed2(ctx, b, outs, ws, class_mask=mask)
D.check(ctx, ws)
ascii_ = D.unpack_ascii(ctx, b, n).cpu().numpy()
off = b.offsets[: n + 1].cpu().numpy().astype(np.uint64)
want = oracle.canonicalize_batch(ascii_, off, normalize=False, threads=8)
arena = outs.out.cpu().numpy()
starts = ctx.aligned_starts(off).astype(np.int64)
lens = (off[1:] - off[:-1]).astype(np.int64)
gs = outs.start[:n].cpu().numpy().astype(np.uint32); gt = outs.strand[:n].cpu().numpy(); gh = outs.hash[:n].cpu().numpy().astype(np.uint64)
bad_start = np.nonzero(gs != want["start"])[0]; bad_strand = np.nonzero(gt != want["strand"])[0]; bad_hash = np.nonzero(gh != want["hash"])[0]
bad_out = []
for i in range(n):
    o, L = int(off[i]), int(lens[i])
    if not np.array_equal(arena[starts[i]: starts[i] + L], want["out"][o:o + L]):
        bad_out.append(i)
print("n", n, "bad start", len(bad_start), "bad strand", len(bad_strand), "bad hash", len(bad_hash), "bad out", len(bad_out))
ok_pos = np.setdiff1d(np.arange(n), np.union1d(bad_start, bad_strand))
bo = np.intersect1d(ok_pos, np.array(bad_out, dtype=np.int64))
bh = np.intersect1d(ok_pos, bad_hash)
print("with right start+strand: bad out", len(bo), "bad hash", len(bh), " bad hash but good out", len(np.setdiff1d(bh, bo)))
for i in list(bad_start[:6]) + list(bad_strand[:3]):
    print("rec", i, "n", lens[i], "got start/strand", gs[i], gt[i], "want", want["start"][i], want["strand"][i])
for i in bo[:4]:
    o, L = int(off[i]), int(lens[i])
    g = arena[starts[i]: starts[i] + L]; w = want["out"][o:o + L]
    d = np.nonzero(g != w)[0]
    print("out rec", i, "n", L, "strand", gt[i], "start", gs[i], "first diff at", d[:5], "ndiff", len(d))
    print("  got ", g[:48].tobytes(), "\n  want", w[:48].tobytes())
    k = int(d[0]) // 16 * 16
    print("  at", k, g[k:k+32].tobytes(), w[k:k+32].tobytes())
for i in np.setdiff1d(bh, bo)[:4]:
    print("hash rec", i, "n", lens[i], hex(gh[i]), hex(want["hash"][i]))
