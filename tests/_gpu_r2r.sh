# run R: 1-CTA kernel back on warp-uniform producer / MMA loops: timeline-stamp library at full size, quick parity subset, one bench line
set +e
mkdir -p gpurun_out
T="I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_trace.so I2IT_TRACE=1"
( env $T timeout 200 python tests/ncu_target.py 8 512 ) > gpurun_out/r2r_trace_stdout.log 2> gpurun_out/r2r_gemm_timeline_trace.txt
( env $T timeout 200 python tests/ncu_target.py 8 512 ) > gpurun_out/r2r_trace_stdout2.log 2> gpurun_out/r2r_gemm_timeline_trace2.txt
( timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -m gpu -q -x 2>&1 | tail -3 ) > gpurun_out/r2r_pytest.log 2>&1
( timeout 300 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline --profile-out gpurun_out/r2r_launch_table.json ) > gpurun_out/r2r_bench.log 2>&1
cat gpurun_out/r2r_trace_stdout.log gpurun_out/r2r_trace_stdout2.log; grep -c TRACE gpurun_out/r2r_gemm_timeline_trace.txt gpurun_out/r2r_gemm_timeline_trace2.txt; grep -h "Error" gpurun_out/r2r_gemm_timeline_trace*.txt | head -3
cat gpurun_out/r2r_pytest.log | cut -c1-200
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r2r_bench.log") if l.startswith("{")][-1]
print(round(d["ms_per_step"],2),"ms", round(d["value"],1),"img/s e2e",round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"])
PY
