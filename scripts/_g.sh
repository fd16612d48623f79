mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25) > gpurun_out/r2_t4.log 2>&1; tail -25 gpurun_out/r2_t4.log | cut -c1-250
export FW_BENCH_SKIP_CPU=1
timeout 300 python bench.py --only c2,c3,c4 --steps 20 --warmup 5 2>gpurun_out/r2_b1.err > gpurun_out/r2_b1.json
python - <<'P'
import json
for f in ("r2_b1",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f,d["ms_per_step"],d["roofline"]["frac"],d["parity"],d["gpu_launches"])
        for k,v in d.get("configs",{}).items(): print("   ",k,v["ms_per_step"],v["roofline"]["frac"],v["roofline"]["kernel_ms"],v["clocks"]["sm_mhz"],v["parity"])
    except Exception as e: print(f,"ERR",e, open(f"gpurun_out/{f}.err").read()[-1500:])
P
