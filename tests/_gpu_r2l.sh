# run L: flash attention with the row sum taken by the PV product (constant ones row, N = 80) and FMNMX3 maxima; sanitizer passes on the new code
set +e
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py tests/test_gpu_clip.py tests/test_gpu_boundary.py -m gpu -q -x 2>&1 ) > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline"
( $B --profile-out gpurun_out/r2l_launch_table.json ) > gpurun_out/r2l_bench.log 2>&1
( $B ) > gpurun_out/r2l_bench_again.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_target.py > gpurun_out/r2l_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2l_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/sanitize_target.py 1 64 > gpurun_out/r2l_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2l_racecheck.log
NCU="ncu --clock-control none --profile-from-start off"
timeout 600 $NCU --set full --import-source on -k regex:flash_attn -c 1 -o gpurun_out/r2l_flash python tests/ncu_target.py > gpurun_out/r2l_ncu_b.log 2>&1
tail -3 gpurun_out/r2l_pytest.log; tail -4 gpurun_out/r2l_memcheck.log; tail -4 gpurun_out/r2l_racecheck.log
for f in r2l_bench r2l_bench_again; do echo "== $f"; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/$f.log") if l.startswith("{")][-1]
    print(round(d["ms_per_step"],2),"ms", round(d["value"],1),"img/s e2e",round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_median"), d["clocks"]["samples"])
except Exception as ex: print("ERR", ex)
PY
done
ls -la gpurun_out | grep r2l
