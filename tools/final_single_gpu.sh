# Final single-GPU measurement pass of a round: bench lines of the four GPU configs, the ncu launch list of the
# headline command and full captures of the dominant kernels.  Run under gpurun; results land in gpurun_out/.
#   bash tools/final_single_gpu.sh <tag>
tag=${1:-r2z}
mkdir -p gpurun_out
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err
timeout 400 python bench.py --workload c2 --steps 30 --warmup 5 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
timeout 400 python bench.py --workload c4 --steps 8 --warmup 3 > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err
timeout 400 python bench.py --workload c5 --steps 8 --warmup 3 > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_c3_reference_arm.json 2> /dev/null
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
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
# full captures: exported to CSV and summarised on the box (the reports themselves are 15 MB each: too much to bring back)
capture() {   # capture <name> <kernel regex> <skip> <summary filter> <command...>
    local name=$1 regex=$2 skip=$3 filt=$4; shift 4
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c 1 -o gpurun_out/${tag}_$name "$@" > gpurun_out/${tag}_ncu_$name.log 2>&1
    ncu -i gpurun_out/${tag}_$name.ncu-rep --page raw --csv > gpurun_out/${tag}_$name.raw.csv 2>/dev/null
    python tools/summarize_ncu.py raw gpurun_out/${tag}_$name.raw.csv gpurun_out/${tag}_${name}_ncu_full.csv "$filt" > /dev/null 2>&1
    rm -f gpurun_out/${tag}_$name.ncu-rep gpurun_out/${tag}_$name.raw.csv
}
capture nn1_kernel_c3 nn1_kernel 2 nn1_kernel python tools/run_chamfer.py
capture nn1_kernel_c2 nn1_kernel 1 nn1_kernel python tools/run_knn.py 1
capture knn_thread_kernel_k16_c4 knn_thread_kernel 1 knn_thread_kernel python tools/run_knn.py 16 10000000 1000000
capture nn1_kernel_c5 nn1_kernel 1 nn1_kernel python tools/run_chamfer.py 65536 65536 1024
capture kd_build_kernel kd_build_kernel 0 kd_build_kernel python tools/run_knn.py 16
ls -la gpurun_out/${tag}_* | awk '{print $5, $9}'
timeout 120 python tools/host_path_time.py > gpurun_out/${tag}_host_path.log 2>&1; head -12 gpurun_out/${tag}_host_path.log
