(time python -m pytest tests -m gpu -x -q -k "prepared") > gpurun_out/r2m_pytest.log 2>&1
tail -12 gpurun_out/r2m_pytest.log
bash tools/final_single_gpu.sh r2z
