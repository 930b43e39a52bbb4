#!/bin/bash
# Final round-2 evidence pass (one GPU): smoke, full -m gpu suite, default bench line, then the launch lists (codec step, H-Codec-1.5 step,
# LM forward) and one ncu --set full capture of the tcgen05 attention kernel.  Outputs under gpurun_out/.
tag=${1:-r2g}
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?")
tail -3 gpurun_out/${tag}_smoke.log
bash profiles/scripts/run_tests_and_bench.sh ${tag}
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_launches_codec.csv \
    python bench.py --workload codec --no-graph --quick --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_codec.log 2>&1
echo "codec launch list rc=$? lines $(wc -l < gpurun_out/${tag}_launches_codec.csv)"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_h15.csv \
    python bench.py --workload h15 --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_h15.log 2>&1
echo "h15 launch list rc=$? lines $(wc -l < gpurun_out/${tag}_launches_h15.csv)"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_lm_forward.csv \
    python profiles/lm_forward_profile.py > gpurun_out/${tag}_ncu_lmf.log 2>&1
echo "lm forward launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa5_kernel -c 4 -o gpurun_out/${tag}_att_umma -f \
    python profiles/scripts/att_bench.py quick > gpurun_out/${tag}_att_ncu.log 2>&1; echo "ncu full rc=$?"
python profiles/scripts/att_bench.py > gpurun_out/${tag}_att_bench.jsonl 2>/dev/null; cat gpurun_out/${tag}_att_bench.jsonl
