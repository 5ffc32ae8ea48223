# Round-2 evidence run (one B200): tests, bench lines, launch lists, full ncu captures of the three round-2 kernels.
set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_final_gputests.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final_bench_q256.json 2> gpurun_out/r2_final_bench_q256.err
for w in r64 m128 mq128; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-fit > gpurun_out/r2_final_bench_$w.json 2> gpurun_out/r2_final_bench_$w.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 3 --no-fit > gpurun_out/r2_final_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_tcq -s 4 -c 2 -o gpurun_out/r2_final_tcq -f python scripts/prof_q256.py 1048576 1 > gpurun_out/r2_final_tcq.log 2>&1
PROF_PRESET=RealNVP PROF_D=64 timeout 400 ncu --set full --clock-control none --import-source on -k regex:flow_tca -s 4 -c 2 -o gpurun_out/r2_final_tca -f python scripts/prof_q256.py 1048576 1 > gpurun_out/r2_final_tca.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wide_ -s 0 -c 14 -o gpurun_out/r2_final_wide -f python scripts/prof_wide.py 16384 1 > gpurun_out/r2_final_wide.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_wide_launches.csv python scripts/prof_wide.py 16384 1 > gpurun_out/r2_final_wide_launch.log 2>&1
cat gpurun_out/r2_final_gputests.txt
tail -c 300 gpurun_out/r2_final_bench_q256.err
