import sys,json
d=json.loads(sys.stdin.readline()); print(d["value"], d["e2e"]["value"], d.get("stages",{}).get("ms_per_step"), "frac", d.get("roofline",{}).get("frac"))
