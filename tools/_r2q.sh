mkdir -p gpurun_out
python tools/numa_probe.py > gpurun_out/r2q_numa.log 2>&1; cat gpurun_out/r2q_numa.log
python bench.py --steps 30 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/r2q_bench_bind.json 2> gpurun_out/r2q_bench_bind.err
PCU_BENCH_NO_BIND=1 python bench.py --steps 30 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/r2q_bench_nobind.json 2> gpurun_out/r2q_bench_nobind.err
python bench.py --steps 30 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/r2q_bench_bind2.json 2> gpurun_out/r2q_bench_bind2.err
for f in bind nobind bind2; do python - <<PY
import json
d = json.loads(open("gpurun_out/r2q_bench_$f.json").read().strip().splitlines()[-1])
print("$f", "e2e ms", d["e2e"]["ms_per_step"], "pageable", d["e2e"]["pageable"]["ms_per_step"], d["e2e"]["stage_ms"], d["config"].get("host_placement"))
PY
done
python -m pytest tests/test_normals.py -m gpu -x -q 2>&1 | tail -5
