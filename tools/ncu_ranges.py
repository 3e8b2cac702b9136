import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
fname = sys.argv[2]
ranges = [tuple(map(int, a.split("-"))) for a in sys.argv[3:]]
cur=None; hdr=None; per={}
tot=0; tots=0
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r[0]=="Line No": hdr=r; continue
    if r[0] not in ("","Function Name") and hdr:
        d=dict(zip(hdr,r))
        try: i=int(d["Instructions Executed"] or 0); s=int(d["Warp Stall Sampling (All Samples)"] or 0); ln=int(r[0])
        except ValueError: continue
        tot+=i; tots+=s
        key = (cur, ln)
        per[key]=per.get(key,(0,0)); per[key]=(per[key][0]+i, per[key][1]+s)
other_i=0; other_s=0
acc={rg:[0,0] for rg in ranges}
for (f,ln),(i,s) in per.items():
    hit=False
    if f==fname:
        for rg in ranges:
            if rg[0]<=ln<=rg[1]: acc[rg][0]+=i; acc[rg][1]+=s; hit=True; break
    if not hit: other_i+=i; other_s+=s
for rg in ranges: print(f"{fname}:{rg[0]}-{rg[1]}: inst {acc[rg][0]/tot*100:5.1f}%  stall {acc[rg][1]/tots*100:5.1f}%")
print(f"other files/lines: inst {other_i/tot*100:5.1f}% stall {other_s/tots*100:5.1f}%")
