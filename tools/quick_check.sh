# parity tests + hot-loop timings + per-kernel durations of one step (ncu launch list) on the current build
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for w in ${WORKS:-kitti cfg3 cfg2}; do
  for ph in fwd bwd; do echo "== $ph $w: $(timeout 300 python tools/run_hot.py --phase $ph --workload $w --iters 20 2>&1 | tail -1)"; done
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/quick_launches.csv python tools/run_hot.py --phase both --iters 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/quick_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=i;break
H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value')
data=[(r[ki], float(r[vi].replace(',',''))) for r in rows[hdr+2:] if len(r)>vi]
for n,v in data[-19:]: print(f"{v/1000:9.1f} us  {n[:80]}")
PY
