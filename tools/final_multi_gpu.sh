# Scaling check on one box: bash tools/final_multi_gpu.sh <tag> <N...>   (under gpurun --gpus 8)
tag=${1:-r2z}; shift
mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = "1" ]; then
    timeout 400 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_scale_n1.json 2> gpurun_out/${tag}_scale_n1.err
  else
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/${tag}_scale_n$n.json 2> gpurun_out/${tag}_scale_n$n.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_scale_n$n.json").read().strip().splitlines()[-1])
    c5 = d.get("c5_strong") or {}
    print("N=$n: value %.3e (%.4f ms)  e2e %.3e (%.3f ms)  c5_strong %.3e (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], c5.get("value", 0), c5.get("ms_per_step", 0)))
except Exception as e:
    print("N=$n failed:", e)
PY
done
