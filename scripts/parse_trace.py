"""Parse a `[ba dbg] tag:delta ...` clock64 trace (BA_DBG_TIMES=1, -DBA_PANEL_TRACE) of the panel kernel: per panel,
how long each phase of the staging groups and of mma warp 0 took."""
import sys
line = [l for l in sys.stdin if l.startswith("[ba dbg]")][-1]
t, ev = 0, {}
for tok in line.split()[2:]:
    tag, d = tok.split(":")
    t += int(d)
    ev[int(tag)] = t
names = {1: "group barrier", 2: "issue", 3: "mbar wait", 4: "empty wait", 5: "bookkeeping", 6: "obs loop"}
for n in range(200, 208):
    st = [ev.get(10 * n + k) for k in range(1, 7)]
    mm = [ev.get(1000 + 10 * n + k) for k in range(1, 4)]
    steps = [tg - 5000 for tg in ev if 5000 <= tg < 6000 and ev[tg] > (mm[2] or 0) - 1 and ev[tg] < (mm[2] or 0) + 400]
    s = f"panel {n}: staging " + " ".join(f"{names[k + 1]} {st[k] - st[k - 1]}" for k in range(1, 6) if st[k] and st[k - 1])
    if ev.get(10 * n + 7) and st[0]:
        s += f" [issue: to expect_tx {ev[10 * n + 7] - st[0]}, expect_tx {ev[10 * n + 8] - ev[10 * n + 7]}, copies {st[1] - ev[10 * n + 8]}]"
    if all(mm):
        s += f" | mma: masks {mm[1] - mm[0]}, tile loop {mm[2] - mm[1]}" + (f" ({steps[0]} steps)" if steps else "")
    print(s)

# all eight mma warps: start (after the full barrier) and end of the strip loop, relative to warp 0's start
for k in range(8):
    row = []
    t0 = ev.get(20000 + 10 * k + 1)
    for mw in range(8):
        a, b = ev.get(20000 + 1000 * mw + 10 * k + 1), ev.get(20000 + 1000 * mw + 10 * k + 3)
        if a and b and t0:
            row.append(f"w{mw} {a - t0:+d}..{b - t0}")
    if row:
        print(f"panel {200 + k}: mma warps " + "  ".join(row))
