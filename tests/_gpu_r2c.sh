set +e
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s 2>&1 ) > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
( time timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r2c_launch_table.json ) > gpurun_out/r2c_bench.log 2>&1
B="timeout 600 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline"
( I2IT_NO_IDRES=1 $B --profile-out gpurun_out/r2c_launch_table_noidres.json ) > gpurun_out/r2c_bench_noidres.log 2>&1
( I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_s8.so $B --profile-out gpurun_out/r2c_launch_table_s8.json ) > gpurun_out/r2c_bench_s8.log 2>&1
( I2IT_NO_GNEPI=1 $B ) > gpurun_out/r2c_bench_nogn.log 2>&1
( $B ) > gpurun_out/r2c_bench_again.log 2>&1
( I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_trace.so I2IT_TRACE=1 timeout 300 python tests/ncu_target.py 8 512 ) > gpurun_out/r2c_trace_stdout.log 2> gpurun_out/r2c_gemm_timeline_trace.txt
NCU="ncu --clock-control none --profile-from-start off"
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2c_ncu_launches.csv python tests/ncu_target.py > gpurun_out/r2c_ncu_launches.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:tapgemm2_kernel -c 14 -o gpurun_out/r2c_prof_tapgemm2 python tests/ncu_target.py > gpurun_out/r2c_ncu_a.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:tapgemm_kernel -s 30 -c 6 -o gpurun_out/r2c_prof_tapgemm python tests/ncu_target.py > gpurun_out/r2c_ncu_b.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"flash_attn|gn_apply|gn_part|gn_stats|layernorm" -c 12 -o gpurun_out/r2c_prof_misc python tests/ncu_target.py > gpurun_out/r2c_ncu_c.log 2>&1
tail -4 gpurun_out/r2c_pytest.log
for f in r2c_bench r2c_bench_noidres r2c_bench_s8 r2c_bench_nogn r2c_bench_again; do echo "== $f"; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/$f.log") if l.startswith("{")][-1]
    print(round(d["ms_per_step"],2),"ms", round(d["value"],1),"img/s e2e",round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_median"))
except Exception as ex: print("ERR", ex)
PY
done
ls -la gpurun_out | grep r2c
