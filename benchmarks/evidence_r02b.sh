set -x
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -c 300 gpurun_out/r02_bench_1gpu.json; tail -3 gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>&1
python bench.py --particles 4736 --steps 2 --warmup 3 --no-cpu-baseline --no-sub > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_p4736.csv python bench.py --particles 4736 --steps 2 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/ncu_launch.log 2>&1
python benchmarks/pgram_once.py 3 | tail -1; ncu --set full --import-source on --clock-control none -k regex:k_abc_pgram -c 1 -o gpurun_out/r02_k_abc_pgram python benchmarks/pgram_once.py 1 > gpurun_out/ncu_pgram.log 2>&1; tail -1 gpurun_out/ncu_pgram.log
python benchmarks/acw_time.py | tail -1
python benchmarks/osc_accuracy.py > gpurun_out/r02_osc_accuracy.log 2>&1; cat gpurun_out/r02_osc_accuracy.log
