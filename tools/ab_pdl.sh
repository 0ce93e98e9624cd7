set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pdl_tests.log
for w in kitti8 cfg2 cfg3; do
  wl="--no-cpu-baseline"; [ $w != kitti8 ] && wl="--workload $w --no-cpu-baseline"
  for m in pdl nopdl; do
    if [ $m = nopdl ]; then export CUSTMA_NO_PDL=1; else unset CUSTMA_NO_PDL; fi
    timeout 300 python bench.py --steps 30 --warmup 5 $wl > gpurun_out/pdl_${w}_${m}.json 2> gpurun_out/pdl_${w}_${m}.err
    timeout 300 python bench.py --steps 30 --warmup 5 $wl --graph > gpurun_out/pdl_${w}_${m}_graph.json 2>> gpurun_out/pdl_${w}_${m}.err
  done
done
unset CUSTMA_NO_PDL
timeout 300 python tools/run_head.py > gpurun_out/pdl_head.log 2>&1
CUSTMA_NO_PDL=1 timeout 300 python tools/run_head.py > gpurun_out/pdl_head_nopdl.log 2>&1
