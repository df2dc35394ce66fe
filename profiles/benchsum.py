import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
print("value",d["value"],"ms/step", d["ms_per_step"],"e2e", d["e2e"]["value"], "decode", d.get("decode",{}).get("value"))
n=int(sys.argv[2]) if len(sys.argv)>2 else 14
print(", ".join("%s %.2f"%(k["kernel"],k["ms_per_step"]) for k in d["kernels"][:n]))
