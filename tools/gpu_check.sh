timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for wl in c4 c5; do
timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_${wl}_b.json 2>gpurun_out/err.txt; tail -2 gpurun_out/err.txt
done
python - <<PY
import json
for f in ("gpurun_out/bench_c4_b.json","gpurun_out/bench_c5_b.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f ms/step %.1f main %.2f TF frac %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"]), {k:round(v["ms_per_step"],2) for k,v in d["phases"].items()})
        print({k:(round(v["iters_per_s"],1), round(v["cholesky_tflops"],2)) for k,v in d.get("hlm",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
timeout 120 python tools/hlm_bench.py 5000 6 | tail -1
timeout 100 python tools/hlm_bench.py 100 2000 | tail -1
timeout 100 python tools/hlm_bench.py 506 300 | tail -1
