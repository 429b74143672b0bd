"""SASS listing of an .ncu-rep with per-instruction executed counts and stall samples.
usage: python tools/ncu_sass.py REP [first_idx] [last_idx]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for k, r in enumerate(data):
    if k < lo or k > hi:
        continue
    s = int(r[ci["# Samples"]])
    st = sorted(((int(r[j] or 0), hdr[j][6:]) for j in stall), reverse=True)[:2]
    sts = " ".join(f"{n}:{v}" for v, n in st if v)
    print(f"{k:5d} {int(r[ci['Instructions Executed']]):>10d} {float(r[ci['Avg. Threads Executed']] or 0):5.1f} {s:6d}  {r[ci['Source']][:70]:70s} {sts}")
