# run H: epilogue instruction diet (magic-number tile decode, alpha*acc+bias as one FFMA with vector bias loads, shifts for the GN indexing) + compile-time LEAN epilogue variant
set +e
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py tests/test_gpu_clip.py tests/test_gpu_boundary.py -m gpu -q -x 2>&1 ) > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline"
( $B --profile-out gpurun_out/r2h_launch_table.json ) > gpurun_out/r2h_bench.log 2>&1
( I2IT_NO_LEAN=1 $B --profile-out gpurun_out/r2h_launch_table_nolean.json ) > gpurun_out/r2h_bench_nolean.log 2>&1
( $B ) > gpurun_out/r2h_bench_again.log 2>&1
( I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_trace.so I2IT_TRACE=1 timeout 300 python tests/ncu_target.py 8 512 ) > gpurun_out/r2h_trace_stdout.log 2> gpurun_out/r2h_gemm_timeline_trace.txt
NCU="ncu --clock-control none --profile-from-start off"
timeout 600 $NCU --set full --import-source on -k regex:tapgemm2_kernel -c 2 -o gpurun_out/r2h_tg2 python tests/ncu_target.py > gpurun_out/r2h_ncu_a.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:flash_attn -c 1 -o gpurun_out/r2h_flash python tests/ncu_target.py > gpurun_out/r2h_ncu_b.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:tapgemm_kernel -s 40 -c 3 -o gpurun_out/r2h_tg1 python tests/ncu_target.py > gpurun_out/r2h_ncu_c.log 2>&1
tail -3 gpurun_out/r2h_pytest.log
for f in r2h_bench r2h_bench_nolean r2h_bench_again; do echo "== $f"; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/$f.log") if l.startswith("{")][-1]
    print(round(d["ms_per_step"],2),"ms", round(d["value"],1),"img/s e2e",round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_median"), d["clocks"]["samples"])
except Exception as ex: print("ERR", ex)
PY
done
ls -la gpurun_out | grep r2h
