mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], 'value %.3e e2e %.3f ms pageable %.3f ms' % (d['value'], d['e2e']['ms_per_step'], d['e2e']['pageable']['ms_per_step']), d['e2e'].get('stage_ms'), (d.get('prepared_target') or {}).get('ms_per_step'), ((d.get('prepared_target') or {}).get('e2e') or {}).get('ms_per_step'))" $1; }
python bench.py --steps 30 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/r2n_direct.json 2>/dev/null; show gpurun_out/r2n_direct.json
OMP_NUM_THREADS=1 python bench.py --steps 30 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/r2n_omp1.json 2>/dev/null; show gpurun_out/r2n_omp1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 1 --steps 30 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/r2n_torchrun1.json 2>/dev/null; show gpurun_out/r2n_torchrun1.json
python tools/host_path_time.py | head -3
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2n_pytest.log 2>&1
tail -5 gpurun_out/r2n_pytest.log
