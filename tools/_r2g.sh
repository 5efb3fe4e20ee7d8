mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2g_pytest.log 2>&1
tail -12 gpurun_out/r2g_pytest.log
export PYTHONFAULTHANDLER=1
python -u -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
echo "torchrun exit $?"
tail -c 1500 gpurun_out/r2g_bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2g_bench_n2.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.4f e2e %.3e (%.3f ms; pageable %.3f ms) launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["ms_per_step"], d["gpu_launches"]))
        print("   stages", d["roofline"]["stage_ms"])
        c5 = d.get("c5_strong"); print("   c5_strong %.3e %.3f ms" % (c5["value"], c5["ms_per_step"]), c5["stage_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_c4.json 2> gpurun_out/r2g_bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_c4.json')); print('c4 value %.3e ms %.3f e2e %.3e (%.1f ms; pageable %.1f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['pageable']['ms_per_step'])); print(d['roofline']['stage_ms'])"
tail -c 300 gpurun_out/r2g_bench_c4.err
