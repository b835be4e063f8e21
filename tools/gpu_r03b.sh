#!/bin/bash
set -u
OUT=gpurun_out/${1:-r03b}
mkdir -p $OUT
export A0_LIB=agent0_b200/libagent0_b200_trace.so
A0_GATHER_WAVES=1,1,2,16 timeout 200 python tools/trace_step.py 32 20 > $OUT/trace_b32_w.txt 2>&1
A0_GATHER_WAVES=auto timeout 200 python tools/trace_step.py 512 20 > $OUT/trace_b512_auto.txt 2>&1
A0_GATHER_WAVES=none timeout 200 python tools/trace_step.py 512 20 > $OUT/trace_b512_none.txt 2>&1
tail -32 $OUT/trace_b32_w.txt; tail -30 $OUT/trace_b512_auto.txt
