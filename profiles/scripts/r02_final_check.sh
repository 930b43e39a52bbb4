#!/bin/bash
# round-2 final check on one B200: lanes A/B, the full -m gpu suite with durations, then the driver's bench invocation (wall time noted)
mkdir -p gpurun_out
t0=$(date +%s)
(timeout 240 python profiles/scripts/lm_lanes_ab.py ${LANES_MODE:-} > gpurun_out/lm_lanes_ab.log 2>&1; echo "lanes rc=$?" >> gpurun_out/lm_lanes_ab.log)
tail -20 gpurun_out/lm_lanes_ab.log
t1=$(date +%s); echo "lanes wall $((t1-t0)) s"
(timeout ${TEST_LIMIT:-900} python -m pytest tests -m gpu -q --durations=25 > gpurun_out/r2h_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_tests.log)
tail -45 gpurun_out/r2h_tests.log
t2=$(date +%s); echo "tests wall $((t2-t1)) s"
(timeout ${BENCH_LIMIT:-420} python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?")
t3=$(date +%s); echo "bench wall $((t3-t2)) s"
tail -c 600 gpurun_out/r2h_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2h_bench.json"))
    print("codec ms", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "gemm frac", d["roofline"]["frac"], "parity ok", (d.get("parity") or {}).get("ok"), "wall", d.get("bench_wall_s"))
    for k,v in (d.get("secondary") or {}).items():
        print(" ", k, v.get("value"), v.get("ms_per_step"), v.get("error"), v.get("skipped"), (v.get("roofline") or {}).get("frac"), v.get("leg_wall_s"))
except Exception as e:
    print("bench json unreadable", e)
PY
