# multi-GPU run 3: what the per-step all-gather costs under different NCCL settings (tests/gpu_gather_diag.py)
set +e
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( $T --master-port 29511 tests/gpu_gather_diag.py ) > gpurun_out/r2e3_diag_default.log 2>&1
( NCCL_MAX_NCHANNELS=1 NCCL_MIN_NCHANNELS=1 $T --master-port 29512 tests/gpu_gather_diag.py ) > gpurun_out/r2e3_diag_1ch.log 2>&1
( NCCL_NVLS_ENABLE=0 NCCL_MAX_NCHANNELS=1 NCCL_MIN_NCHANNELS=1 NCCL_MAX_CTAS=1 $T --master-port 29514 tests/gpu_gather_diag.py ) > gpurun_out/r2e3_diag_1ch_nonvls.log 2>&1
grep -h "^{" gpurun_out/r2e3_diag_*.log
grep -il "error\|Traceback" gpurun_out/r2e3_diag_*.log
