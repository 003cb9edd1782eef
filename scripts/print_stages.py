import json,sys
d=json.load(open(sys.argv[1])); print(round(d["value"],1), round(d["e2e"]["value"],1)); print({k:v["ms_per_call"] for k,v in d["stages"].items()})
