# per-kernel durations of the small kernels of a step (ncu, cold caches) and the device step, for one or more
# library builds: bash tools/ncu_small.sh [lib.so ...]
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active
LIBS="${@:-default}"
for lib in $LIBS; do
  if [ "$lib" != default ]; then export LRVB_LIB_PATH=$lib; else unset LRVB_LIB_PATH; fi
  tag=$(basename $lib .so)
  for w in c3 c2; do
    ncu --metrics $M --clock-control none -k regex:"k_finish|k_prep|k_global_post|k_csr" --launch-skip 30 -c 16 --csv --log-file gpurun_out/r02b_${w}_${tag}.csv python tools/ab_step.py $w 3 > /dev/null 2>&1
  done
  for w in c3 c2; do python tools/ab_step.py $w 20; done
done
