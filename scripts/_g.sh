mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x -k "reverb or conv or config5 or golden" 2>&1 | tail -8) > gpurun_out/r2_t6.log 2>&1; tail -6 gpurun_out/r2_t6.log | cut -c1-300
export FW_BENCH_SKIP_CPU=1
FW_BENCH_C5_VOICES=8192 timeout 200 python bench.py --only c5 --steps 20 --warmup 5 2>gpurun_out/r2_c5_tail.err > gpurun_out/r2_c5_tail.json
timeout 200 python bench.py --only c4 --steps 20 --warmup 5 2>gpurun_out/r2_c4_b.err > gpurun_out/r2_c4_b.json
python - <<'P'
import json
for f in ("r2_c5_tail","r2_c4_b"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f,d["ms_per_step"],d["roofline"]["frac"],d["roofline"]["step_share"]["temporal_ms"],d["clocks"]["sm_mhz"],d["parity"])
    except Exception as e: print(f,"ERR",e, open(f"gpurun_out/{f}.err").read()[-1500:])
P
