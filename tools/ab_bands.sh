for w in cfg2 kitti; do
for nb in 0 3 4 5 6 7 8 9 10 12 13 14; do
  export CUSTMA_BANDS=$nb
  echo "$w bands=$nb $(python tools/run_hot.py --phase fwd --workload $w --iters 20 2>&1 | tail -1 | cut -c1-28) | $(python tools/run_hot.py --phase bwd --workload $w --iters 20 2>&1 | tail -1 | cut -c1-28)"
done; done
