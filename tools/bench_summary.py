import sys,json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith("{"): continue
    d=json.loads(line)
    print(d["value"], [(s["kernel"][:6], round(s["ms"],3)) for s in d.get("stages",[])], d.get("verified_bit_exact_payload"), "e2e", d["e2e"]["value"])
