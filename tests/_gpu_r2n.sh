# run N: final evidence — full GPU test suite, smoke(), the default bench line with all configs + CPU arm, trace-library crash under compute-sanitizer
set +e
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 ) > gpurun_out/r2n_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest_full.log
( python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2n_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2n_smoke.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r2n_launch_table.json ) > gpurun_out/r2n_bench.log 2>&1
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r2n_bench_reference.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline --batch 1 ) > gpurun_out/r2n_bench_b1.log 2>&1
T="I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_trace.so I2IT_TRACE=1"
( env $T timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/ncu_target.py 1 256 2>&1 | grep -v "^TRACE" | head -60 ) > gpurun_out/r2n_trace_memcheck_256.log 2>&1
tail -3 gpurun_out/r2n_pytest_full.log; tail -2 gpurun_out/r2n_smoke.log; tail -12 gpurun_out/r2n_trace_memcheck_256.log | cut -c1-220
for f in r2n_bench r2n_bench_b1; do echo "== $f"; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/$f.log") if l.startswith("{")][-1]
    print(round(d["ms_per_step"],2),"ms", round(d["value"],1),"img/s e2e",round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_median"), d["clocks"]["samples"], d.get("cpu_baseline"))
    print({k:(v.get("value"), v.get("ms_per_step")) if "value" in v else [(s["per_gpu_batch"],round(s["value"],1)) for s in v.get("sweep",[])] for k,v in d["configs"].items()})
except Exception as ex: print("ERR", ex)
PY
done
grep "^{" gpurun_out/r2n_bench_reference.log | cut -c1-400
ls -la gpurun_out | grep r2n
