set +e
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s 2>&1 ) > gpurun_out/r2d_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest_full.log
( I2IT_FLASH_V1=1 I2IT_NO_SPLITK=1 I2IT_NO_CATFUSE=1 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py tests/test_gpu_clip.py -m gpu -q 2>&1 ) > gpurun_out/r2d_pytest_v1_nosplitk.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest_v1_nosplitk.log
B="timeout 600 python bench.py --steps 10 --warmup 3 --configs none --no-cpu-baseline"
( time timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r2d_launch_table.json ) > gpurun_out/r2d_bench.log 2>&1
( $B --profile-out gpurun_out/r2d_launch_table_b.json ) > gpurun_out/r2d_bench_again.log 2>&1
( I2IT_NO_IDRES=1 $B --profile-out gpurun_out/r2d_launch_table_noidres.json ) > gpurun_out/r2d_bench_noidres.log 2>&1
( I2IT_NO_GNEPI=1 $B ) > gpurun_out/r2d_bench_nogn.log 2>&1
( I2IT_FLASH_V1=1 $B --profile-out gpurun_out/r2d_launch_table_flashv1.json ) > gpurun_out/r2d_bench_flashv1.log 2>&1
( I2IT_NO_SPLITK=1 $B ) > gpurun_out/r2d_bench_nosplitk.log 2>&1
( I2IT_NO_CATFUSE=1 $B ) > gpurun_out/r2d_bench_nocatfuse.log 2>&1
( I2IT_NO_HALO=1 $B ) > gpurun_out/r2d_bench_nohalo.log 2>&1
( $B --batch 1 ) > gpurun_out/r2d_bench_b1.log 2>&1
( I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_trace.so I2IT_TRACE=1 timeout 300 python tests/ncu_target.py 8 512 ) > gpurun_out/r2d_trace_stdout.log 2> gpurun_out/r2d_gemm_timeline_trace.txt
NCU="ncu --clock-control none --profile-from-start off"
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2d_ncu_launches.csv python tests/ncu_target.py > gpurun_out/r2d_ncu_launches.log 2>&1
mkdir -p /tmp/ncu
timeout 600 $NCU --set full --import-source on -k regex:tapgemm2_kernel -c 8 -o /tmp/ncu/tg2 python tests/ncu_target.py > gpurun_out/r2d_ncu_a.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:tapgemm_kernel -s 30 -c 6 -o /tmp/ncu/tg1 python tests/ncu_target.py > gpurun_out/r2d_ncu_b.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"flash_attn|gn_apply|gn_part|gn_stats|layernorm" -c 12 -o /tmp/ncu/misc python tests/ncu_target.py > gpurun_out/r2d_ncu_c.log 2>&1
for n in tg2 tg1 misc; do
  python profiles/summarize_ncu.py /tmp/ncu/$n.ncu-rep gpurun_out/r2d_${n}_ncu_full.csv >> gpurun_out/r2d_ncu_sum.log 2>&1
  ncu -i /tmp/ncu/$n.ncu-rep --page source --csv > /tmp/ncu/$n.src.csv 2>/dev/null
  python profiles/line_stalls.py /tmp/ncu/$n.src.csv > gpurun_out/r2d_${n}_line_stalls.txt 2>> gpurun_out/r2d_ncu_sum.log
done
tail -3 gpurun_out/r2d_pytest_full.log; tail -3 gpurun_out/r2d_pytest_v1_nosplitk.log
for f in r2d_bench r2d_bench_again r2d_bench_noidres r2d_bench_nogn r2d_bench_flashv1 r2d_bench_nosplitk r2d_bench_nocatfuse r2d_bench_nohalo r2d_bench_b1; do echo "== $f"; python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/$f.log") if l.startswith("{")][-1]
    print(round(d["ms_per_step"],2),"ms", round(d["value"],1),"img/s e2e",round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_median"), d["clocks"]["samples"])
except Exception as ex: print("ERR", ex)
PY
done
du -sh gpurun_out; ls -la gpurun_out | grep r2d
