"""One-line summary of bench.py JSON lines: python tools/bench_summary.py [file ...]  (stdin when no file is given)."""
import json
import sys

files = [open(f) for f in sys.argv[1:]] or [sys.stdin]
for fh in files:
    for line in fh:
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        print(d["value"], [(s["kernel"][:6], round(s["ms"], 3)) for s in d.get("stages", [])], d.get("verified_bit_exact_payload"),
              "e2e", d["e2e"]["value"], "e2e_sc16", d.get("e2e_sc16", {}).get("value"))
