"""Top stall sites of one kernel from an ncu report's source page:  python tools/ncu_hot.py report.ncu-rep <kernel-id> [n]"""
import csv, subprocess, sys, io
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:100])
h = rows[1]
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
body = [r for r in rows[2:] if len(r) > isamp and r[isamp].isdigit()]
tot = sum(int(r[isamp] or 0) for r in body)
print("total samples", tot)
order = sorted(range(len(body)), key=lambda i: -int(body[i][isamp] or 0))[:n]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[c] or 0), h[c]) for c in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%%  ex %8s  %-70s %s" % (i, 100.0 * int(r[isamp]) / max(tot, 1), r[iex], r[isrc].strip()[:70], ", ".join("%s=%d" % (b, a) for a, b in st if a)))
