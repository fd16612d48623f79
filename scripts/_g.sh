mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4) > gpurun_out/r2_t8.log 2>&1; tail -2 gpurun_out/r2_t8.log | cut -c1-250
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2_full.err > gpurun_out/r2_full.json
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_full.json").read().strip().splitlines()[-1]); print(d["ms_per_step"],d["roofline"]["frac"],d["clocks"])
for k,v in d.get("configs",{}).items(): print("   ",k,v["ms_per_step"],v["roofline"]["frac"],v["roofline"]["kernel_ms"],v["clocks"]["sm_mhz"],v["parity"].get("ok"), v.get("tflops"))
P
