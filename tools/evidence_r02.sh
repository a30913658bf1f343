# Final evidence of round 2 on ONE B200 (gpurun): plain bench lines first, profiler passes after them.
#   bash tools/evidence_r02.sh bench     plain bench lines + launch lists      (outputs under gpurun_out/r02f_*)
#   bash tools/evidence_r02.sh ncu       --set full captures, summarised on the box (gpurun brings back <= 64 MiB)
O=gpurun_out
if [ "$1" = bench ]; then
python bench.py --impl reference > $O/r02f_bench_reference.json 2> $O/r02f_bench_reference.err
python bench.py > $O/r02f_bench_c2c3.json 2> $O/r02f_bench_c2c3.err
python bench.py --workload c1 > $O/r02f_bench_c1.json 2> $O/r02f_bench_c1.err
python bench.py --workload c4 > $O/r02f_bench_c4.json 2> $O/r02f_bench_c4.err
python bench.py --workload c5 > $O/r02f_bench_c5.json 2> $O/r02f_bench_c5.err
# launch list of bench.py itself (C2 only: --no-target), per-launch durations
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02f_launches_bench_c2.csv \
    python bench.py --steps 3 --warmup 3 --no-target > $O/r02f_ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file $O/r02f_launches_c4s.csv python tools/profile_eval.py c4s 2 > /dev/null 2>&1
else
# --set full captures of the dominant kernels (a later evaluation of each driver); the reports are summarised
# here and only the summaries (+ the small C2 report) travel back
ncu --set full --clock-control none --import-source on -k regex:"k_fused_eval" --launch-skip 2 -c 1 \
    -o $O/r02f_prof_c2 python tools/profile_eval.py c2 3 > $O/r02f_ncu_c2.log 2>&1
ncu --set full --clock-control none -k regex:"k_finish|k_csr_pass" --launch-skip 4 -c 2 \
    -o $O/r02f_prof_c2tail python tools/profile_eval.py c2 3 >> $O/r02f_ncu_c2.log 2>&1
ncu --set full --clock-control none -k regex:"k_obs_fused|k_gram_mid" --launch-skip 2 -c 2 \
    -o $O/r02f_prof_c3 python tools/profile_eval.py c3 2 > $O/r02f_ncu_c3.log 2>&1
ncu --set full --clock-control none -k regex:"k_obs|k_group|k_gram_wide" --launch-skip 3 -c 3 \
    -o $O/r02f_prof_c4s python tools/profile_eval.py c4s 2 > $O/r02f_ncu_c4s.log 2>&1
python tools/summarize_profiles.py ncu $O/r02f_ncu_c2_summary.md "r02 final, ncu --set full, C2 (N=1M K=20 G=10k): one-pass kernel, finish kernel, CSR refill (python tools/profile_eval.py c2)" $O/r02f_prof_c2.ncu-rep $O/r02f_prof_c2tail.ncu-rep
python tools/summarize_profiles.py ncu $O/r02f_ncu_c3_summary.md "r02 final, ncu --set full, C3 (N=10M K=50 G=100k): observation pass and team Gram kernel (python tools/profile_eval.py c3)" $O/r02f_prof_c3.ncu-rep
python tools/summarize_profiles.py ncu $O/r02f_ncu_c4s_summary.md "r02 final, ncu --set full, configs[3] slice (N=500k K=200 G=5k): k_obs, k_group, k_gram_wide (python tools/profile_eval.py c4s)" $O/r02f_prof_c4s.ncu-rep
python - <<'PY'
import json, sys
sys.path.insert(0, "tools")
import summarize_profiles as sp
out = {}
for tag, rep in (("c2", "gpurun_out/r02f_prof_c2.ncu-rep"), ("c3", "gpurun_out/r02f_prof_c3.ncu-rep"), ("c4s", "gpurun_out/r02f_prof_c4s.ncu-rep")):
    for d in sp.raw(rep):
        def b(key):
            v, u = d[key]
            return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        out.setdefault(tag, {})[d["Kernel Name"][0].split("(")[0]] = int(b("dram__bytes_read.sum") + b("dram__bytes_write.sum"))
json.dump(out, open("gpurun_out/r02f_dram_bytes.json", "w"), indent=1)
print(out)
PY
rm -f $O/r02f_prof_c2tail.ncu-rep $O/r02f_prof_c3.ncu-rep $O/r02f_prof_c4s.ncu-rep
fi
ls -la $O/r02f_*
