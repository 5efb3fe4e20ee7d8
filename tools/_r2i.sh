mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q -k "kd_replica or tie_replay or knn_matches_reference_goldens or knn_against_oracle or million or topk or c4_full or normals") > gpurun_out/r2i_pytest.log 2>&1
tail -12 gpurun_out/r2i_pytest.log
python tools/replay_time.py > gpurun_out/r2i_replay.log 2>&1
cat gpurun_out/r2i_replay.log
