"""Summarise a k_post2_tc clock trace written by scripts/gpu_trace.py (run here, on the saved text)."""
import sys
txt = open(sys.argv[1]).read().splitlines()
ev = {}
for l in txt:
    if l.startswith(("rc", "kernel")):
        print(l)
    for name in ("MMA", "A", "B"):
        if l.startswith(name + " "):
            ev[name] = dict(tuple(map(int, x.split(":"))) for x in l.split()[1:])
mma, A, B = ev["MMA"], ev["A"], ev["B"]
print("tile: MMA out start / ctx ready / issued | wait_y start / ready || E1: start, out_full, stats, abar, packed, y_full arrive, o' stored")
for t in range(4):
    print(t, [mma.get(k + t) for k in (1000, 1100, 1200, 2000, 2100)], [A.get(k + t) for k in (1000, 1100, 1120, 1130, 1140, 1200, 1300)])
print("chunk: F2 wait start, ready | E2 wait start, acch ready, done | E2 cycles, done->F2 ready, period, issuer waited")
prev = None
for i in range(32):
    m3, m4 = mma.get(3000 + i), mma.get(4000 + i)
    a3, a4, a5 = A.get(3000 + i), A.get(4000 + i), A.get(5000 + i)
    if m4 is None:
        break
    print(i, m3, m4, "|", a3, a4, a5, "|", a5 - a4, m4 - a5, (m4 - prev) if prev else None, m4 - m3)
    prev = m4
print("E3:", [(k, B[k]) for k in sorted(B)])
