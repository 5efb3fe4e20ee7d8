# Final single-GPU measurement pass of a round: bench lines of the four GPU configs, the ncu launch list of the
# headline command and full captures of the dominant kernels.  Run under gpurun; results land in gpurun_out/.
#   bash tools/final_single_gpu.sh <tag>
tag=${1:-r2z}
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err
python bench.py --workload c2 --steps 30 --warmup 5 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
python bench.py --workload c4 --steps 8 --warmup 3 > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err
python bench.py --workload c5 --steps 8 --warmup 3 > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_c3_reference_arm.json 2> /dev/null
for w in c3 c2 c4 c5; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_${w}.json").read().strip().splitlines()[-1])
    print("${w}: value %.3e  %.4f ms/step  e2e %.3e (%.3f ms)  roofline frac %.4f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"]))
    print("    ", d["roofline"]["stage_ms"])
except Exception as e:
    print("${w} failed:", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'nn1_kernel' -s 2 -c 1 -o gpurun_out/${tag}_nn1 python tools/run_chamfer.py > gpurun_out/${tag}_ncu_nn1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'nn1_kernel' -s 1 -c 1 -o gpurun_out/${tag}_nn1_c2 python tools/run_knn.py 1 > gpurun_out/${tag}_ncu_nn1_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'knn_thread_kernel' -s 1 -c 1 -o gpurun_out/${tag}_knn16_c4 python tools/run_knn.py 16 10000000 1000000 > gpurun_out/${tag}_ncu_knn16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'nn1_kernel' -s 1 -c 1 -o gpurun_out/${tag}_nn1_c5 python tools/run_chamfer.py 65536 65536 1024 > gpurun_out/${tag}_ncu_nn1_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'kd_build_kernel' -s 0 -c 1 -o gpurun_out/${tag}_kdbuild python tools/run_knn.py 16 > gpurun_out/${tag}_ncu_kdbuild.log 2>&1
ls -la gpurun_out/${tag}_* | awk '{print $5, $9}'
python tools/host_path_time.py > gpurun_out/${tag}_host_path.log 2>&1; head -12 gpurun_out/${tag}_host_path.log
