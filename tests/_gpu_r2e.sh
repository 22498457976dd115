# multi-GPU run (gpurun --gpus N): two-rank parity + scaling diagnostics.  N from the first argument (2 or 8).
set +e
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,power.limit --format=csv > gpurun_out/r2e_n${N}_smi.txt 2>&1
if [ "$N" = "2" ]; then
  ( python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/gpu_two_rank.py ) > gpurun_out/r2e_two_rank_tiny.log 2>&1
  echo "rc=$?" >> gpurun_out/r2e_two_rank_tiny.log
  ( I2IT_TWO_RANK_FULL=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tests/gpu_two_rank.py ) > gpurun_out/r2e_two_rank_full.log 2>&1
  echo "rc=$?" >> gpurun_out/r2e_two_rank_full.log
  ( python -m pytest tests/test_gpu_boundary.py -m gpu -q -k two_rank ) > gpurun_out/r2e_two_rank_pytest.log 2>&1
  grep -h "TWO_RANK_OK\|rc=\|passed\|failed" gpurun_out/r2e_two_rank_*.log
fi
( time NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $N --steps 20 --warmup 5 --configs 4 ) > gpurun_out/r2e_bench_n${N}.log 2>&1
( python bench.py --gpus 1 --steps 20 --warmup 5 --configs none --no-cpu-baseline ) > gpurun_out/r2e_bench_n1_samebox.log 2>&1
python - <<PY
import json
for f in ("gpurun_out/r2e_bench_n${N}.log", "gpurun_out/r2e_bench_n1_samebox.log"):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, d["n_gpus"], round(d["ms_per_step"],2), "ms", round(d["value"],1), "img/s e2e", round(d["e2e"]["value"],1), d.get("scaling_diag"), {k:(v.get("value"), v.get("ms_per_step")) for k,v in d["configs"].items() if isinstance(v,dict)})
    except Exception as ex: print(f, "ERR", ex)
PY
grep -c "NCCL INFO" gpurun_out/r2e_bench_n${N}.log; grep -m3 "NVLS\|nranks" gpurun_out/r2e_bench_n${N}.log
