set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -c 600 gpurun_out/r02_bench_1gpu.json; tail -3 gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>&1; tail -c 400 gpurun_out/r02_bench_reference_arm.json
python bench.py --particles 4736 --steps 2 --warmup 3 --no-cpu-baseline --no-sub > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_p4736.csv python bench.py --particles 4736 --steps 2 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_abc_acf_tc -c 1 -o gpurun_out/r02_k_abc_acf_tc_fullsize python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-sub > gpurun_out/ncu_tc.log 2>&1; tail -2 gpurun_out/ncu_tc.log
NP=592 ncu --set full --import-source on --clock-control none -k regex:k_abc_psd2 -c 1 -o gpurun_out/r02_k_abc_psd2 python benchmarks/psd_once.py 1 > gpurun_out/ncu_psd2.log 2>&1; tail -1 gpurun_out/ncu_psd2.log
ncu --set full --import-source on --clock-control none -k regex:k_acw_stream -c 1 -o gpurun_out/r02_k_acw_stream python benchmarks/acw_once.py 1 > gpurun_out/ncu_acw.log 2>&1; tail -1 gpurun_out/ncu_acw.log
python benchmarks/pgram_once.py 3; ncu --set full --import-source on --clock-control none -k regex:k_abc_pgram -c 1 -o gpurun_out/r02_k_abc_pgram python benchmarks/pgram_once.py 1 > gpurun_out/ncu_pgram.log 2>&1; tail -1 gpurun_out/ncu_pgram.log
