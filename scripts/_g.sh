mkdir -p gpurun_out
export FW_BENCH_SKIP_CPU=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --only c2 --steps 20 --warmup 5 2>gpurun_out/r2_n8c2.err > gpurun_out/r2_n8c2.json
for i in 0 3 6; do CUDA_VISIBLE_DEVICES=$i timeout 100 python bench.py --only c2 --steps 20 --warmup 5 --min-seconds 0.3 2>/dev/null > gpurun_out/r2_gpu$i.json; done
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_n8c2.json").read().strip().splitlines()[-1]); print("n8",d["value"],d["ms_per_step"],d["ms_per_step_by_rank"],d.get("bus_parity"))
for i in (0,3,6):
    try:
        d=json.loads(open(f"gpurun_out/r2_gpu{i}.json").read().strip().splitlines()[-1]); print("gpu",i,d["ms_per_step"],d["roofline"]["kernel_ms"])
    except Exception as e: print(i,"ERR",e)
P
