import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"launches",d.get("gpu_launches"))
print("e2e",json.dumps(d["e2e"])[:600])
print("day",json.dumps(d.get("day"),indent=1)[:3500])
print({k:(round(v["avg_ms"],4),round(v.get("frac",0),3)) for k,v in d["kernels"].items() if v["share"]>0.005})
print("roofline",json.dumps(d["roofline"])[:500])
print(d.get("cpu_baseline")); print(d.get("clocks"))
