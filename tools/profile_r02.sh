#!/bin/bash
# ncu captures of round 2 (run under gpurun; summaries are written to profiles/ by tools/ncu_summary.py afterwards)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD1="python bench.py --samples 8 --steps 2 --warmup 3 --no-cpu --no-hlm"
$CMD1 > gpurun_out/plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/r02_launches_c5.csv $CMD1 > gpurun_out/ncu1.log 2>&1
echo "launch list c5 rc=$?"
$CMD1 > gpurun_out/plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gradient_tma -s 3 -c 1 -f -o gpurun_out/r02_prof_main_c5 $CMD1 > gpurun_out/ncu2.log 2>&1
echo "full main rc=$?"
# ncu replays kernels one at a time: the look-ahead schedule (kernels that wait for each other) cannot run under it
export SPW_LA=0
CMD2="python tools/hlm_bench.py 5000 2 matern2 1"
$CMD2 > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/r02_launches_hlm5000.csv $CMD2 > gpurun_out/ncu3.log 2>&1
echo "launch list hlm rc=$?"
export SPW_DAG=1
$CMD2 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:potrf_dag -s 3 -c 1 -f -o gpurun_out/r02_prof_dag_hlm5000 $CMD2 > gpurun_out/ncu4.log 2>&1
echo "full dag rc=$?"
tail -2 gpurun_out/plain2.log gpurun_out/plain3.log
ls -la gpurun_out | tail -12
