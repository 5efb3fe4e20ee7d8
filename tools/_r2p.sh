mkdir -p gpurun_out
python -m pytest tests/test_normals.py tests/test_dedup.py -m gpu -x -q > gpurun_out/r2p_tests_new.log 2>&1; tail -15 gpurun_out/r2p_tests_new.log
python tools/next_rows_time.py > gpurun_out/r2p_rows.log 2>&1; cat gpurun_out/r2p_rows.log
python bench.py --steps 30 --warmup 3 > gpurun_out/r2p_bench_c3.json 2> gpurun_out/r2p_bench_c3.err; tail -c 3000 gpurun_out/r2p_bench_c3.json
ncu --set full --clock-control none --import-source on -k regex:'nn1_kernel' -s 2 -c 1 -o gpurun_out/r2p_nn1 python tools/run_chamfer.py > gpurun_out/r2p_ncu_nn1.log 2>&1; tail -3 gpurun_out/r2p_ncu_nn1.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2p_tests_all.log 2>&1; tail -5 gpurun_out/r2p_tests_all.log
