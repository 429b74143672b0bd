"""Per-source-line view of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py REP [top_n] [file_filter]
Prints, for the CUDA source lines with the most executed warp-instructions / stall samples:
instructions, share, samples, share, avg active threads, and the dominant stall reasons."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
flt = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        names = r
        continue
    if hdr is None or r[0] == "":
        continue
    lines.append((cur_file, r))
ci = hdr
stall_cols = [i for i, h in enumerate(names) if h.startswith("stall_") and "Not Issued" not in h]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


tot_i = sum(num(r[ci["Instructions Executed"]]) for _, r in lines)
tot_s = sum(num(r[ci["# Samples"]]) for _, r in lines)
print(f"total warp-instructions {tot_i}, samples {tot_s}")
sel = [(f, r) for f, r in lines if flt in f]
sel.sort(key=lambda fr: -num(fr[1][ci["Instructions Executed"]]))
print("file:line  inst%  samp%  thr  top stalls | source")
for f, r in sel[:top]:
    i = num(r[ci["Instructions Executed"]])
    s = num(r[ci["# Samples"]])
    th = num(r[ci["Thread Instructions Executed"]])
    st = sorted(((num(r[k]), names[k][6:]) for k in stall_cols), reverse=True)[:3]
    sts = " ".join(f"{n}:{100 * v // max(s, 1)}%" for v, n in st if v)
    print(f"{f}:{r[0]:>4} {100 * i / tot_i:5.1f} {100 * s / max(tot_s, 1):5.1f} {th / max(i, 1):5.1f}  {sts} | {r[1].strip()[:90]}")
