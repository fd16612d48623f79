mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6) > gpurun_out/r2_t5.log 2>&1; tail -3 gpurun_out/r2_t5.log | cut -c1-250
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2_full.err > gpurun_out/r2_full.json
python - <<'P'
import json
for f in ("r2_full",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f,d["ms_per_step"],d["roofline"]["frac"],d["clocks"], d.get("passes"), d["e2e"]["value"])
        for k,v in d.get("configs",{}).items(): print("   ",k,v["ms_per_step"],v["roofline"]["frac"],v["roofline"]["kernel_ms"],v["clocks"]["sm_mhz"],v["parity"], round(v["bench_wall_s"],1), v.get("tflops"), v.get("block_sized_calls"))
    except Exception as e: print(f,"ERR",e, open(f"gpurun_out/{f}.err").read()[-1500:])
P
export FW_BENCH_SKIP_CPU=1
for w in c2 c3 c4 dag; do timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_$w.csv python bench.py --only $w --steps 2 --warmup 3 --min-seconds 0.0001 > gpurun_out/ncu_l_$w.log 2>&1; done
FW_BENCH_C5_VOICES=8192 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c5_8192.csv python bench.py --only c5 --steps 2 --warmup 3 --min-seconds 0.0001 > gpurun_out/ncu_l_c5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 12 -c 1 -o gpurun_out/r02_c2_chain python bench.py --only c2 --steps 2 --warmup 3 --min-seconds 0.0001 > gpurun_out/ncu_f_c2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:biquad_delay_lanes -s 8 -c 1 -o gpurun_out/r02_c3_temporal python bench.py --only c3 --steps 2 --warmup 3 --min-seconds 0.0001 > gpurun_out/ncu_f_c3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:reverb_gemm2 -s 8 -c 1 -o gpurun_out/r02_c4_gemm2 python bench.py --only c4 --steps 2 --warmup 3 --min-seconds 0.0001 > gpurun_out/ncu_f_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_* | tail -12
