set +e
mkdir -p gpurun_out
T="I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_trace.so I2IT_TRACE=1"
( env $T timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/ncu_target.py 8 512 2>&1 | grep -v "^TRACE" | head -70 ) > gpurun_out/r2o_trace_memcheck_512.log 2>&1
tail -40 gpurun_out/r2o_trace_memcheck_512.log | cut -c1-250
