# Round-2 evidence run (one B200): tests, bench lines, launch list, full ncu capture of the MADE spline kernel.
# (The tcq / tca / wide captures under profiles/ were taken with the same kernel sources: collect_profiles_kernels.sh.)
set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_final_gputests.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final_bench_q256.json 2> gpurun_out/r2_final_bench_q256.err
for w in r64 m128 mq128; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-fit > gpurun_out/r2_final_bench_$w.json 2> gpurun_out/r2_final_bench_$w.err; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 3 --no-fit > gpurun_out/r2_final_ncu_launch.log 2>&1
PROF_PRESET=MaskedAutoregressiveRQNSF PROF_D=128 timeout 400 ncu --set full --clock-control none --import-source on -k regex:flow_tcm -s 2 -c 1 -o gpurun_out/r2_final_tcm -f python scripts/prof_q256.py 262144 1 > gpurun_out/r2_final_tcm.log 2>&1
cat gpurun_out/r2_final_gputests.txt
tail -c 300 gpurun_out/r2_final_bench_q256.err
