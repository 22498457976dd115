set +e
mkdir -p gpurun_out
L="I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_trace.so"
( env $L I2IT_SYNC_EACH=1 timeout 200 python tests/ncu_target.py 8 512 2>&1 | grep -v "^TRACE" | tail -3 ) > gpurun_out/r2q_tracelib_notrace.log 2>&1
( env $L I2IT_TRACE=1 I2IT_SYNC_EACH=1 I2IT_NO_TMAOUT=1 timeout 200 python tests/ncu_target.py 8 512 2>&1 | grep -v "^TRACE" | tail -3 ) > gpurun_out/r2q_trace_notmaout.log 2>&1
for f in tracelib_notrace trace_notmaout; do echo "== $f"; cat gpurun_out/r2q_$f.log | cut -c1-300; done
