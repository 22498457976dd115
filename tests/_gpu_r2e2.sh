# multi-GPU run 2 (gpurun --gpus N): all-gather overlapped with the next step (I2IT_OVERLAP_GATHER=1) vs synchronous (default), same box
set +e
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( $T --master-port 29541 tests/gpu_two_rank.py ) > gpurun_out/r2e2_two_rank_tiny.log 2>&1
echo "rc=$?" >> gpurun_out/r2e2_two_rank_tiny.log
( time NCCL_DEBUG=INFO $T --master-port 29500 bench.py --gpus $N --steps 20 --warmup 5 --configs 4 ) > gpurun_out/r2e2_bench_n${N}.log 2>&1
( I2IT_OVERLAP_GATHER=1 $T --master-port 29501 bench.py --gpus $N --steps 20 --warmup 5 --configs none --no-cpu-baseline ) > gpurun_out/r2e2_bench_n${N}_sync.log 2>&1
( $T --master-port 29502 bench.py --gpus $N --steps 20 --warmup 5 --configs none --no-cpu-baseline ) > gpurun_out/r2e2_bench_n${N}_again.log 2>&1
( python bench.py --gpus 1 --steps 20 --warmup 5 --configs none --no-cpu-baseline ) > gpurun_out/r2e2_bench_n1_samebox.log 2>&1
python - <<PY
import json
for f in ("gpurun_out/r2e2_bench_n${N}.log", "gpurun_out/r2e2_bench_n${N}_sync.log", "gpurun_out/r2e2_bench_n${N}_again.log", "gpurun_out/r2e2_bench_n1_samebox.log"):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, d["n_gpus"], round(d["ms_per_step"],2), "ms", round(d["value"],1), "img/s e2e", round(d["e2e"]["value"],1), d["config"].get("collective"), d.get("scaling_diag"), {k:(v.get("value"), v.get("ms_per_step")) for k,v in d["configs"].items() if isinstance(v,dict)})
    except Exception as ex: print(f, "ERR", ex)
PY
grep -h "TWO_RANK_OK\|rc=" gpurun_out/r2e2_two_rank_tiny.log; grep -c "NCCL INFO" gpurun_out/r2e2_bench_n${N}.log; grep -m3 "NVLS\|nranks" gpurun_out/r2e2_bench_n${N}.log
grep -i "error\|Traceback" gpurun_out/r2e2_bench_n${N}.log | head -5
