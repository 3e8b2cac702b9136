"""Summarise an ncu source-page CSV (--print-source cuda,sass) per source line: instructions and stall samples."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
lines = []
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] != "" and hdr:
        d = dict(zip(hdr, r))
        try:
            lines.append((cur_file, int(r[0]), r[1].strip()[:90], int(d["Instructions Executed"] or 0), int(d["Warp Stall Sampling (All Samples)"] or 0), d))
        except ValueError:
            pass
tot_i = sum(l[3] for l in lines); tot_s = sum(l[4] for l in lines)
print(f"total instr {tot_i}  samples {tot_s}")
print("--- by instructions")
for f, n, s, i, st, d in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{f}:{n:4d} {i/tot_i*100:5.1f}% inst {st/max(tot_s,1)*100:5.1f}% stall | {s}")
print("--- by stall samples")
for f, n, s, i, st, d in sorted(lines, key=lambda l: -l[4])[:top]:
    keys = [k for k in d if k.startswith("stall_") and "Not Issued" not in k and d[k] not in ("", "0")]
    keys = sorted(keys, key=lambda k: -int(d[k]))[:3]
    print(f"{f}:{n:4d} {st/max(tot_s,1)*100:5.1f}% stall {i/tot_i*100:5.1f}% inst | {s[:60]} | " + " ".join(f"{k[6:]}={d[k]}" for k in keys))
