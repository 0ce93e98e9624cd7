timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for w in kitti cfg2 cfg3; do
  for m in prepare noprepare; do
    fl=""; [ $m = noprepare ] && fl="--no-prepare"
    python bench.py --steps 30 --warmup 5 --workload $w --no-cpu-baseline --no-e2e $fl > gpurun_out/prep_${m}_$w.json 2> gpurun_out/prep_${m}_$w.err
    python bench.py --steps 30 --warmup 5 --workload $w --no-cpu-baseline --no-e2e --graph $fl > gpurun_out/prep_${m}_${w}_graph.json 2>> gpurun_out/prep_${m}_$w.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/prep_${m}_$w.json").read().strip().splitlines()[-1]); g=json.loads(open("gpurun_out/prep_${m}_${w}_graph.json").read().strip().splitlines()[-1])
print("$m $w step", round(d["ms_per_step"],4), "graph", round(g["ms_per_step"],4), d["gpu_launches"])
PY
  done
done
