#!/bin/bash
# full -m gpu test suite, then the default bench line (the driver's invocation); logs under gpurun_out/
tag=${1:-r2}
mkdir -p gpurun_out
(timeout 1300 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_tests.log)
tail -4 gpurun_out/${tag}_tests.log
(timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?")
tail -c 400 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("codec ms", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "gemm frac", d["roofline"]["frac"], "parity ok", (d.get("parity") or {}).get("ok"))
for k,v in (d.get("secondary") or {}).items():
    print(" ", k, v.get("value"), v.get("ms_per_step"), v.get("error"), (v.get("roofline") or {}).get("frac"))
PY
