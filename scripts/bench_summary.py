import json,sys
d=json.loads(sys.stdin.read())
print("obs/s %.4g  step %.4f ms  kernels %s  conv dev %.3f ms api %.2f ms it %d" % (d["value"], d["ms_per_step"], {k:round(v,4) for k,v in d["kernels_ms"].items()}, d["ba_converge"]["device_ms"], (d["ba_converge"]["api"] or {}).get("wall_ms",0), d["ba_converge"]["iterations"]))
