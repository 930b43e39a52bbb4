#!/bin/bash
# attention kernels: event-timed comparison, then one ncu --set full capture of the tcgen05 kernels
mkdir -p gpurun_out
timeout 300 python profiles/scripts/att_bench.py > gpurun_out/att_bench.jsonl 2> gpurun_out/att_bench.err; echo "bench rc=$?"; cat gpurun_out/att_bench.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa5_ -c 10 -o gpurun_out/att_umma -f python profiles/scripts/att_bench.py quick > gpurun_out/att_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/att_umma.ncu-rep
