python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
python tools/knn_occupancy.py 2>&1 | tee gpurun_out/r2t_knn_occ.log | head -26
python tools/nonuniform_time.py 2>&1 | tail -20
