M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for w in c3 c2; do
ncu --metrics $M --clock-control none -k regex:"k_finish|k_prep|k_global_post|k_csr" --launch-skip 30 -c 16 --csv --log-file gpurun_out/r02b_${w}_new2.csv python tools/ab_step.py $w 3 > /dev/null 2>&1
done
for w in c3 c2; do python tools/ab_step.py $w 20; done
