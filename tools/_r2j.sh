mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q -k "kd_replica or tie_replay") > gpurun_out/r2j_pytest.log 2>&1
tail -6 gpurun_out/r2j_pytest.log
python tools/replay_time.py > gpurun_out/r2j_replay.log 2>&1
cat gpurun_out/r2j_replay.log
