# run F: epilogue rework (vectorised GN statistics, pipelined TMEM loads, bias staged once) + last-block GroupNorm reduction
set +e
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x 2>&1 ) > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline"
( $B --profile-out gpurun_out/r2f_launch_table.json ) > gpurun_out/r2f_bench.log 2>&1
( I2IT_NO_GNEPI=1 $B --profile-out gpurun_out/r2f_launch_table_nogn.json ) > gpurun_out/r2f_bench_nogn.log 2>&1
( $B ) > gpurun_out/r2f_bench_again.log 2>&1
( I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_trace.so I2IT_TRACE=1 timeout 300 python tests/ncu_target.py 8 512 ) > gpurun_out/r2f_trace_stdout.log 2> gpurun_out/r2f_gemm_timeline_trace.txt
NCU="ncu --clock-control none --profile-from-start off"
mkdir -p /tmp/ncu
timeout 600 $NCU --set full --import-source on -k regex:tapgemm2_kernel -c 8 -o /tmp/ncu/tg2 python tests/ncu_target.py > gpurun_out/r2f_ncu_a.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"gn_apply|gn_part|gn_stats" -c 12 -o /tmp/ncu/misc python tests/ncu_target.py > gpurun_out/r2f_ncu_c.log 2>&1
for n in tg2 misc; do
  python profiles/summarize_ncu.py /tmp/ncu/$n.ncu-rep gpurun_out/r2f_${n}_ncu_full.csv >> gpurun_out/r2f_ncu_sum.log 2>&1
  ncu -i /tmp/ncu/$n.ncu-rep --page source --csv > /tmp/ncu/$n.src.csv 2>/dev/null
  python profiles/line_stalls.py /tmp/ncu/$n.src.csv > gpurun_out/r2f_${n}_line_stalls.txt 2>> gpurun_out/r2f_ncu_sum.log
done
tail -3 gpurun_out/r2f_pytest.log
for f in r2f_bench r2f_bench_nogn r2f_bench_again; do echo "== $f"; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/$f.log") if l.startswith("{")][-1]
    print(round(d["ms_per_step"],2),"ms", round(d["value"],1),"img/s e2e",round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_median"), d["clocks"]["samples"])
except Exception as ex: print("ERR", ex)
PY
done
ls -la gpurun_out | grep r2f
