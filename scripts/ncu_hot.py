"""Summarise an `ncu --page source --csv` dump: stall mix, hottest instructions, samples per code region."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
base = int(data[0][0], 16)
S = lambda r, k: int(r[ci[k]] or 0)
tot = sum(S(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data), "executed", sum(S(r, "Instructions Executed") for r in data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
mix = {h: sum(S(r, h) for r in data) for h in stalls}
print("stall mix:", {k: v for k, v in sorted(mix.items(), key=lambda x: -x[1]) if v})
reg = collections.Counter(); ins = collections.Counter()
G = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0x2000
for r in data:
    off = int(r[0], 16) - base
    reg[off // G] += S(r, "# Samples"); ins[off // G] += S(r, "Instructions Executed")
for k in sorted(reg):
    if reg[k] > tot * 0.01: print(hex(k * G), "samples", reg[k], "instr executed", ins[k])
for r in sorted(data, key=lambda r: -S(r, "# Samples"))[:45]:
    st = {h[6:]: S(r, h) for h in stalls if S(r, h)}
    print(hex(int(r[0], 16) - base), r[1].strip()[:58].ljust(58), S(r, "# Samples"), S(r, "Instructions Executed"), st)
