for m in single slots; do
  [ $m = single ] && unset CUSTMA_HOST_PIPELINE
  [ $m = slots ] && export CUSTMA_HOST_PIPELINE=slots
  for w in kitti cfg2 cfg3; do
    timeout 300 python bench.py --steps 40 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/host_${m}_$w.json 2> gpurun_out/host_${m}_$w.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/host_${m}_$w.json").read().strip().splitlines()[-1])
print("$m $w device", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "copy-only", round(d["e2e"]["copy_only_ms_per_step"],4), d["e2e"]["matches_device_path"])
PY
  done
  timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "host" 2>&1 | tail -2
done
unset CUSTMA_HOST_PIPELINE
for c in 1 2 4 8; do echo "chunk $c: $(CUSTMA_HOST_CHUNK=$c python tools/e2e_probe2.py 2>&1 | grep sync)"; done
