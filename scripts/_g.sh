mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2_t2.log 2>&1; tail -3 gpurun_out/r2_t2.log
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2_full.err > gpurun_out/r2_full.json
python - <<'P'
import json
for f in ("r2_full",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f,d["ms_per_step"],d["roofline"]["frac"],d["roofline"]["kernel_ms"],d["clocks"], d.get("passes"), d["e2e"]["value"], d["cpu_baseline"])
        for k,v in d.get("configs",{}).items(): print("   ",k,v["ms_per_step"],v["roofline"]["frac"],v["roofline"]["kernel_ms"],v["clocks"]["sm_mhz"],v["parity"], v["bench_wall_s"], v.get("tflops"))
    except Exception as e: print(f,"ERR",e, open(f"gpurun_out/{f}.err").read()[-1500:])
P
