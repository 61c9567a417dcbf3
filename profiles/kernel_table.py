import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
print("  " + " | ".join(f"{k} {1e3*v['ms']/v['launches']:.1f}us x{v['launches']}" for k,v in list(d["kernels"].items())[:7]))
