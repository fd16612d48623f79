mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -q -x -k "plugin" 2>&1 | tail -12) > gpurun_out/r2_t3.log 2>&1; tail -8 gpurun_out/r2_t3.log
(timeout 300 python -m pytest tests -m gpu -q -x -k "reverb or conv or config5 or golden or generic" 2>&1 | tail -8) > gpurun_out/r2_t2.log 2>&1; tail -6 gpurun_out/r2_t2.log
export FW_BENCH_SKIP_CPU=1
timeout 200 python bench.py --only c4 --steps 20 --warmup 5 2>gpurun_out/r2_c4_2cta.err > gpurun_out/r2_c4_2cta.json
FW_REVERB_BN=256 timeout 200 python bench.py --only c4 --steps 20 --warmup 5 2>gpurun_out/r2_c4_2cta256.err > gpurun_out/r2_c4_2cta256.json
FW_REVERB_1CTA=1 timeout 200 python bench.py --only c4 --steps 20 --warmup 5 2>gpurun_out/r2_c4_1cta.err > gpurun_out/r2_c4_1cta.json
FW_BENCH_C5_VOICES=8192 timeout 200 python bench.py --only c5 --steps 20 --warmup 5 2>gpurun_out/r2_c5_2cta.err > gpurun_out/r2_c5_2cta.json
python - <<'P'
import json
for f in ("r2_c4_2cta","r2_c4_2cta256","r2_c4_1cta","r2_c5_2cta"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f,d["ms_per_step"],d["roofline"]["frac"],d["roofline"]["step_share"]["temporal_ms"],d["clocks"]["sm_mhz"],d["parity"])
    except Exception as e: print(f,"ERR",e, open(f"gpurun_out/{f}.err").read()[-1500:])
P
