"""usage: python tools/ncu_sass.py src.csv file.cuh LO HI  -> SASS of source lines LO..HI with executed counts (ncu --page source --print-source cuda,sass --csv)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
fname, lo, hi = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
cur, hdr, line, src = None, None, None, None
tot = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0] == "Function Name" or hdr is None:
        continue
    d = dict(zip(hdr, r))
    if r[0] != "":
        line, src = int(r[0]), r[1].strip()
        if cur == fname and lo <= line <= hi:
            print(f"--- {line}: {src[:110]}")
        continue
    if cur == fname and line is not None and lo <= line <= hi:
        ie = d.get("Instructions Executed") or "0"; ie = int(ie) if ie.isdigit() else 0
        tot += ie
        print(f"      {ie:9d} {d.get('Avg. Threads Executed', ''):>5} {r[3].strip()}")
print("total", tot)
