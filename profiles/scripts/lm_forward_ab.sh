#!/bin/bash
# LM teacher-forced forward: tcgen05 causal attention (default) vs the round-1 mma.sync kernel (QB_ATTENTION=legacy), same box
mkdir -p gpurun_out
for mode in umma legacy umma legacy; do
  QB_ATTENTION=$mode timeout 300 python bench.py --workload lm_forward --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$mode', d['ms_per_step'], d['clocks']['sm_mhz'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lm_forward_launches.csv python profiles/lm_forward_profile.py > gpurun_out/lm_forward_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/lm_forward_launches.csv 0 | head -30
