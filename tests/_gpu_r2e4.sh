set +e
mkdir -p gpurun_out
( I2IT_DIAG_PER_STEP=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/gpu_gather_diag.py ) > gpurun_out/r2e4_diag_per_step.log 2>&1
grep -h "^{" gpurun_out/r2e4_diag_per_step.log; grep -i "Traceback" -A6 gpurun_out/r2e4_diag_per_step.log | head -20
