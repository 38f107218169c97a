python -m pytest tests/test_pool_maxsim_merge_gpu.py tests/test_wrapper_integration.py tests/test_tc_gpu.py -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --workload C4p > gpurun_out/r2q_bench_c4p_n1.json 2> gpurun_out/r2q_c4p.err; tail -2 gpurun_out/r2q_c4p.err; cut -c1-400 gpurun_out/r2q_bench_c4p_n1.json
python bench.py --workload C4 > gpurun_out/r2q_bench_c4_n1.json 2> gpurun_out/r2q_c4.err; tail -2 gpurun_out/r2q_c4.err; cut -c1-300 gpurun_out/r2q_bench_c4_n1.json
