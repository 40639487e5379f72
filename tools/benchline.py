import sys,json
d=json.loads(sys.stdin.readline()); print(d["value"], d["e2e"]["value"], d["stages"]["ms_per_step"])
