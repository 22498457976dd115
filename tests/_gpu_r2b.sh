set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2b_smi.txt 2>&1
# pass 1: round-1 epilogue (per-thread stores, gn_stats kernel) -> validates everything except the new epilogue
( time I2IT_NO_TMAOUT=1 I2IT_NO_GNEPI=1 python -m pytest tests -m gpu -q -s 2>&1 ) > gpurun_out/r2b_pytest_old_epilogue.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest_old_epilogue.log
# pass 2: default (TMA-store epilogue + GroupNorm statistics in the epilogue)
( time python -m pytest tests -m gpu -q -s 2>&1 ) > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
# pass 3: TMA stores without the fused statistics (only the kernel cases + tiny e2e), to separate the two if pass 2 fails
( time I2IT_NO_GNEPI=1 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -m gpu -q -x 2>&1 ) > gpurun_out/r2b_pytest_tma_nogn.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest_tma_nogn.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2b_smoke.log 2>&1
( time timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r2b_launch_table.json ) > gpurun_out/r2b_bench.log 2>&1
( time I2IT_NO_TMAOUT=1 I2IT_NO_GNEPI=1 timeout 600 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline --profile-out gpurun_out/r2b_launch_table_old_epilogue.json ) > gpurun_out/r2b_bench_old_epilogue.log 2>&1
( time I2IT_NO_GNEPI=1 timeout 600 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline --profile-out gpurun_out/r2b_launch_table_tma_nogn.json ) > gpurun_out/r2b_bench_tma_nogn.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_target.py > gpurun_out/r2b_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2b_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/sanitize_target.py 1 64 > gpurun_out/r2b_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2b_racecheck.log
for f in r2b_pytest_old_epilogue r2b_pytest r2b_pytest_tma_nogn; do echo "== $f"; tail -4 gpurun_out/$f.log; done
tail -2 gpurun_out/r2b_smoke.log
for f in r2b_bench r2b_bench_old_epilogue r2b_bench_tma_nogn; do echo "== $f"; tail -c 300 gpurun_out/$f.log; done
tail -3 gpurun_out/r2b_memcheck.log; tail -3 gpurun_out/r2b_racecheck.log
