# run J: single-thread unrolled TMA producer loops, single-thread MMA loop in the 1-CTA kernel, branch-free tile decode
set +e
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py tests/test_gpu_clip.py tests/test_gpu_boundary.py -m gpu -q -x 2>&1 ) > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline"
( $B --profile-out gpurun_out/r2j_launch_table.json ) > gpurun_out/r2j_bench.log 2>&1
( $B ) > gpurun_out/r2j_bench_again.log 2>&1
( I2IT_NO_HALO=1 $B ) > gpurun_out/r2j_bench_nohalo.log 2>&1
( I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_trace.so I2IT_TRACE=1 timeout 300 python tests/ncu_target.py 8 512 ) > gpurun_out/r2j_trace_stdout.log 2> gpurun_out/r2j_gemm_timeline_trace.txt
NCU="ncu --clock-control none --profile-from-start off"
timeout 600 $NCU --set full --import-source on -k regex:tapgemm2_kernel -s 1 -c 2 -o gpurun_out/r2j_tg2 python tests/ncu_target.py > gpurun_out/r2j_ncu_a.log 2>&1
tail -3 gpurun_out/r2j_pytest.log
for f in r2j_bench r2j_bench_again r2j_bench_nohalo; do echo "== $f"; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/$f.log") if l.startswith("{")][-1]
    print(round(d["ms_per_step"],2),"ms", round(d["value"],1),"img/s e2e",round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_median"), d["clocks"]["samples"])
except Exception as ex: print("ERR", ex)
PY
done
ls -la gpurun_out | grep r2j
