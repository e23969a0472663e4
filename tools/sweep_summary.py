import json,sys
for f in sys.argv[1:]:
    rows=[json.loads(l) for l in open(f) if l.startswith("{")]
    seen=set(); out=[]
    for r in rows:
        if r["launch"] in seen: continue
        seen.add(r["launch"]); out.append(r)
    out.sort(key=lambda r:r["ms"])
    print("==",f)
    for r in out[:8]: print(r["ms"], r["gbps"], r["ok"], r["launch"][13:])
    print([ (r["ms"],r["launch"][13:]) for r in rows if r["tile"]==0 and r["ctas"]==0 and r["stages"]==0])
