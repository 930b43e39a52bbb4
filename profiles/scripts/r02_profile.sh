#!/bin/bash
# Round-2 profiling pass (one GPU): launch lists of one codec step / one LM generate, then ncu --set full of the dominant GEMM.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_codec.csv \
    python bench.py --workload codec --no-graph --quick --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_codec.log 2>&1
echo "codec launch list rc=$? lines $(wc -l < gpurun_out/r02_launches_codec.csv)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_lm.csv \
    python profiles/lm_profile.py 8 > gpurun_out/r02_ncu_lm.log 2>&1
echo "lm launch list rc=$? lines $(wc -l < gpurun_out/r02_launches_lm.csv)"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 4 -c 2 -f -o gpurun_out/r02_prof_gemm \
    python profiles/gemm_microbench.py 2 > gpurun_out/r02_ncu_gemm.log 2>&1
echo "gemm full rc=$?"
ls -la gpurun_out/r02_prof_gemm.ncu-rep
