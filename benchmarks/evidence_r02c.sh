# final evidence of round 2 (HEAD): smoke, GPU tests, bench (1 GPU) + reference arm, launch list, ncu captures of the
# acw streaming pass, acw call timings
set -x
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -c 300 gpurun_out/r02_bench_1gpu.json; tail -3 gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>&1
python bench.py --particles 4736 --steps 2 --warmup 3 --no-cpu-baseline --no-sub > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_p4736.csv python bench.py --particles 4736 --steps 2 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_acw_stream -c 1 -o gpurun_out/r02_k_acw_stream python benchmarks/acw_once.py 1 > gpurun_out/ncu_acw.log 2>&1; tail -1 gpurun_out/ncu_acw.log
(python benchmarks/acw_window.py; python benchmarks/acw_host_time.py; python benchmarks/acw_e2e_time.py; python benchmarks/knee_time.py; python benchmarks/lomb_time.py; python benchmarks/pmc_weights_time.py) > gpurun_out/r02_acw_timings.log 2>&1; cat gpurun_out/r02_acw_timings.log
