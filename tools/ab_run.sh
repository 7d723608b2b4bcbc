#!/bin/bash
# Development aid (run on the GPU box): time the conv kernel of several ab/lib_NAME.so builds inside one call.
# usage: tools/ab_run.sh OUTDIR NAME...   (a NAME prefixed with "+" first runs the conv / strip / frame-group parity tests on it)
out=$1; shift; mkdir -p $out
for v in "$@" ; do
  n=${v#+}
  if [ "$v" != "$n" ]; then
    SRCNN_B200_LIB=ab/lib_$n.so timeout ${PTMO:-600} python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "conv_tensor or strips or frame_groups or batch or c2_full or butterfly" 2>&1 | tail -3 | sed "s/^/[$n] /" >> $out/log.txt
  fi
  for rep in ${REPS:-1 2}; do
    SRCNN_B200_LIB=ab/lib_$n.so timeout ${TMO:-45} python tools/conv_time.py 30 2>&1 | tail -1 >> $out/log.txt || echo "$n: FAILED/TIMEOUT" >> $out/log.txt
  done
done
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader >> $out/log.txt
