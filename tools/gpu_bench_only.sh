#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/bo_bench.json 2> gpurun_out/bo_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bo_bench.err | cut -c1-200
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bo_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "clocks", d.get("clocks"))
print("api", {k: (round(v.get("value",0),2) if "value" in v else v) for k,v in (d.get("api") or {}).items() if isinstance(v,dict)})
for k,v in (d.get("configs") or {}).items(): print(k, v.get("parity"), round(v.get("value",0),2))
PY
