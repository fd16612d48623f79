mkdir -p gpurun_out
export FW_BENCH_SKIP_CPU=1
(timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -6) > gpurun_out/r2_multi.log 2>&1; tail -3 gpurun_out/r2_multi.log | cut -c1-300
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --only c2 --steps 20 --warmup 5 2>gpurun_out/r2_n8_$name.err > gpurun_out/r2_n8_$name.json; }
run sig A=1
run ch2 NCCL_MAX_NCHANNELS=2
run ch1 NCCL_MAX_NCHANNELS=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 8 --only c5 --steps 20 --warmup 5 2>gpurun_out/r2_n8_c5.err > gpurun_out/r2_n8_c5.json
python - <<'P'
import json
for f in ("r2_n8_sig","r2_n8_ch2","r2_n8_ch1","r2_n8_c5"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f,d["value"],d["ms_per_step"],d.get("ms_per_step_p10"),d.get("ms_per_step_p90"),d["roofline"]["frac"],d.get("bus_parity"),d.get("bus_identical_on_all_ranks"))
    except Exception as e: print(f,"ERR",e, open(f"gpurun_out/{f}.err").read()[-2500:])
P
