set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
( time python -m pytest tests -m gpu -x -q -s 2>&1 ) > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2a_smoke.log 2>&1
( time timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r2a_launch_table.json ) > gpurun_out/r2a_bench.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_target.py > gpurun_out/r2a_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2a_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/sanitize_target.py 1 64 > gpurun_out/r2a_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2a_racecheck.log
tail -3 gpurun_out/r2a_pytest.log; tail -2 gpurun_out/r2a_smoke.log; tail -c 600 gpurun_out/r2a_bench.log; tail -3 gpurun_out/r2a_memcheck.log; tail -3 gpurun_out/r2a_racecheck.log
