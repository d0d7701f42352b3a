set -x
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01b.csv python tools/step_profile.py > gpurun_out/ncu_step_r01b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_.*tc_kernel -c 4 -o gpurun_out/prof_attn_tc python tools/one_attn.py > gpurun_out/ncu_attn_tc.log 2>&1
ncu -i gpurun_out/prof_attn_tc.ncu-rep --page raw --csv > gpurun_out/prof_attn_tc_raw.csv 2>/dev/null
python bench.py > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; tail -c 600 gpurun_out/bench_r01b.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r01b.json 2>> gpurun_out/bench_r01b.err; tail -c 400 gpurun_out/bench_ref_r01b.json
