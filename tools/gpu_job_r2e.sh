set -x
cd profiles/probes && (timeout 300 ./_bin/tma_gather_probe > tma_gather_probe.out 2>&1; echo "probe rc=$?"; cp tma_gather_probe.out ../../gpurun_out/; tail -20 tma_gather_probe.out); cd ../..
for tool in memcheck initcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r2e_sanitizer_${tool}_smoke.log 2>&1; echo "sanitizer $tool smoke rc=$?"; tail -3 gpurun_out/r2e_sanitizer_${tool}_smoke.log
done
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "souporcell3_three_runs or cell_comm_single_rank" > gpurun_out/r2e_sanitizer_${tool}_s3.log 2>&1; echo "sanitizer $tool s3 rc=$?"; tail -3 gpurun_out/r2e_sanitizer_${tool}_s3.log
done
python bench.py --steps 2 --warmup 1 --no-sub --no-cpu-baseline > gpurun_out/r2e_bench_short.json 2> gpurun_out/r2e_bench_short.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r2e.csv python bench.py --steps 1 --warmup 0 --no-sub --no-cpu-baseline > gpurun_out/r2e_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'estep_kernel|mstep_kernel' --launch-skip 8 -c 2 -o gpurun_out/prof_r2e -f python tools/prof_ncu.py 0 50 12 > gpurun_out/r2e_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
