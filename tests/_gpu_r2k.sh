# run K: identity-residual slab off by default (A/B: I2IT_IDRES=1), three halo stages (lib variant), halo tiles for dual-source convs (lib variant)
set +e
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py tests/test_gpu_clip.py tests/test_gpu_boundary.py -m gpu -q -x 2>&1 ) > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline"
( $B --profile-out gpurun_out/r2k_launch_table.json ) > gpurun_out/r2k_bench.log 2>&1
( I2IT_IDRES=1 $B ) > gpurun_out/r2k_bench_idres.log 2>&1
( I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_halo3.so $B --profile-out gpurun_out/r2k_launch_table_halo3.json ) > gpurun_out/r2k_bench_halo3.log 2>&1
( I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_halox2.so timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -m gpu -q -x 2>&1 ) > gpurun_out/r2k_pytest_halox2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest_halox2.log
( I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_halox2.so $B --profile-out gpurun_out/r2k_launch_table_halox2.json ) > gpurun_out/r2k_bench_halox2.log 2>&1
( $B ) > gpurun_out/r2k_bench_again.log 2>&1
tail -3 gpurun_out/r2k_pytest.log; tail -3 gpurun_out/r2k_pytest_halox2.log
for f in r2k_bench r2k_bench_idres r2k_bench_halo3 r2k_bench_halox2 r2k_bench_again; do echo "== $f"; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/$f.log") if l.startswith("{")][-1]
    print(round(d["ms_per_step"],2),"ms", round(d["value"],1),"img/s e2e",round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_median"), d["clocks"]["samples"])
except Exception as ex: print("ERR", ex)
PY
done
ls -la gpurun_out | grep r2k
