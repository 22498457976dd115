# run M: trace-build crash under compute-sanitizer, PDL A/B, fresh ncu launch list / captures of the current build
set +e
mkdir -p gpurun_out
T="I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_trace.so I2IT_TRACE=1"
( env $T timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_target.py 2>&1 | grep -v "^TRACE" | head -80 ) > gpurun_out/r2m_trace_memcheck.log 2>&1
( env $T timeout 300 python tests/ncu_target.py 8 512 ) > gpurun_out/r2m_trace_stdout.log 2> gpurun_out/r2m_gemm_timeline_trace.txt
B="timeout 600 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline"
( $B ) > gpurun_out/r2m_bench.log 2>&1
( I2IT_PDL=1 $B ) > gpurun_out/r2m_bench_pdl.log 2>&1
( $B ) > gpurun_out/r2m_bench_again.log 2>&1
( I2IT_PDL=1 $B ) > gpurun_out/r2m_bench_pdl_again.log 2>&1
NCU="ncu --clock-control none --profile-from-start off"
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2m_ncu_launches.csv python tests/ncu_target.py > gpurun_out/r2m_ncu_launches.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:tapgemm2_kernel -c 14 -o gpurun_out/r2m_tg2 python tests/ncu_target.py > gpurun_out/r2m_ncu_a.log 2>&1
tail -5 gpurun_out/r2m_trace_memcheck.log; head -3 gpurun_out/r2m_gemm_timeline_trace.txt | cut -c1-200
for f in r2m_bench r2m_bench_pdl r2m_bench_again r2m_bench_pdl_again; do echo "== $f"; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/$f.log") if l.startswith("{")][-1]
    print(round(d["ms_per_step"],2),"ms", round(d["value"],1),"img/s e2e",round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_median"), d["clocks"]["samples"])
except Exception as ex: print("ERR", ex)
PY
done
ls -la gpurun_out | grep r2m
