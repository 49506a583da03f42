set -u
cd "$GRAFT_REPO_ROOT"
CMD1="python bench.py --samples 8 --steps 2 --warmup 3 --no-cpu --no-hlm"
$CMD1 > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv --log-file gpurun_out/r02_launches_c5.csv $CMD1 > gpurun_out/ncu1.log 2>&1
echo "launch list c5 rc=$?"
export SPW_LA=0
CMD2="python tools/hlm_bench.py 5000 2 matern2 1"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/r02_launches_hlm5000.csv $CMD2 > gpurun_out/ncu3.log 2>&1
echo "launch list hlm rc=$?"
