#!/bin/bash
# A/B: parked (suspend-hint) vs spinning barrier waits of the GEMM's epilogue / producer warps; alternate to average out clock drift
NP=unified_audio_b200/lib_nopark/libquark_b200_nopark.so
for i in 1 2; do
  echo "== spin"; QB_LIB=$NP python profiles/gemm_microbench.py 10 2>&1 | grep -E "pwconv|swiglu" | cut -c1-160
  echo "== park"; python profiles/gemm_microbench.py 10 2>&1 | grep -E "pwconv|swiglu" | cut -c1-160
done
for i in 1 2; do
  echo "== step spin"; QB_LIB=$NP python bench.py --workload codec --quick --steps 5 --warmup 3 2>&1 | tail -1
  echo "== step park"; python bench.py --workload codec --quick --steps 5 --warmup 3 2>&1 | tail -1
done
