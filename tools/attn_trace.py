"""Timeline of the whole-row attention kernel (debug build with -DIIC_ATTN_TRACE): prints per warp / tile clock deltas of CTA 0."""
import sys
rows = [list(map(int, l.split())) for l in open(sys.argv[1])]
t0 = min(v for r in rows for v in r[2:] if v > 0)
names = {3: "issuer", 4: "sm q0 qt0", 7: "sm q3 qt0", 8: "sm q0 qt1", 16: "sm q0 qt3"}
for w, nm in names.items():
    print(f"--- warp {w} ({nm})")
    for r in rows:
        if r[0] == w and 4 <= r[1] < 12 and any(r[2:]):
            st = r[2:]
            base = st[0] - t0
            print(f"  tile {r[1]:2d} @ {base:8d}: " + " ".join(f"{(b - a) if (a > 0 and b > 0) else -1:6d}" for a, b in zip(st, st[1:])))
