set +e
mkdir -p gpurun_out
T="I2IT_LIB=$PWD/img2img-turbo_b200/lib/libi2it_trace.so I2IT_TRACE=1"
( env $T I2IT_SYNC_EACH=1 timeout 300 python tests/ncu_target.py 8 512 2>&1 | grep -v "^TRACE" | tail -12 ) > gpurun_out/r2p_trace_sync_each.log 2>&1
( env $T timeout 300 python tests/ncu_target.py 8 512 2>&1 | grep -v "^TRACE" | tail -6 ) > gpurun_out/r2p_trace_plain.log 2>&1
( env $T I2IT_NO_OSTG2=1 timeout 300 python tests/ncu_target.py 8 512 2>&1 | grep -v "^TRACE" | tail -4 ) > gpurun_out/r2p_trace_noostg2.log 2>&1
( env $T I2IT_NO_LEAN=1 timeout 300 python tests/ncu_target.py 8 512 2>&1 | grep -v "^TRACE" | tail -4 ) > gpurun_out/r2p_trace_nolean.log 2>&1
for f in sync_each plain noostg2 nolean; do echo "== $f"; cat gpurun_out/r2p_trace_$f.log | cut -c1-300; done
