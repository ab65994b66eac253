"""Format the PB_TC_PROF=2 event timeline of K2 (stderr lines '[PB_TC_TRACE] tile i role k:cycles ...'): per-role deltas."""
import sys
rows = {}
for line in sys.stdin:
    if "[PB_TC_TRACE]" not in line:
        continue
    f = line.split()
    tile, role = int(f[2]), f[3]
    ev = {int(a.split(":")[0]): int(a.split(":")[1]) for a in f[4:]}
    rows[(tile, role)] = ev
tiles = sorted({k[0] for k in rows})
for t in tiles:
    m = rows.get((t, "mma1"), {})
    if not m:
        continue
    base = min(min(ev.values()) for (tt, r), ev in rows.items() if tt == t and ev)
    print(f"tile {t}: base {base}")
    for role in ("cvt0", "cvt1", "mma1", "drainA", "drainB", "mma2", "out"):
        ev = rows.get((t, role), {})
        ks = sorted(ev)
        print(f"  {role:7s}", " ".join(f"{k}:{ev[k] - base}" for k in ks))
    ks = sorted(m)
    print("  mma1 d  ", " ".join(str(m[k] - m[ks[i - 1]]) if i else "-" for i, k in enumerate(ks)))
    ev = rows.get((t, "mma1x"), {})   # MMA1 issuer internals: 4 pairs x 7 points
    for k in range(4):
        pts = [ev.get(k * 8 + i) for i in range(7)]
        if all(v is not None for v in pts):
            print(f"  mma1 pair {k}: start {pts[0] - base}  slice_wait {pts[1]-pts[0]} a1_wait {pts[2]-pts[1]} fence {pts[3]-pts[2]} elect+issueA {pts[4]-pts[3]} issueB {pts[5]-pts[4]} syncwarp {pts[6]-pts[5]}")
    w, iss = {}, {}
    if w and iss:
        print("  wait    ", " ".join(str(w[k] - m[k - 1]) if k - 1 in m and k in w else "-" for k in ks))
        print("  issue   ", " ".join(str(iss[k] - w[k]) if k in w and k in iss else "-" for k in ks))
        print("  commit+ ", " ".join(str(m[k] - iss[k]) if k in iss else "-" for k in ks))
if len(tiles) > 1:
    a, b = rows[(tiles[0], "mma1")], rows[(tiles[-1], "mma1")]
    print("tile period (mma1 chunk 26):", (b[26] - a[26]) / (tiles[-1] - tiles[0]))
