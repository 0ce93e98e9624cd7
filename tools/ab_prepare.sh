timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "host" 2>&1 | tail -2
for m in prepare noprepare; do
  [ $m = noprepare ] && export CUSTMA_HOST_NO_PREPARE=1
  for P in 8 4; do echo "$m P=$P: $(python tools/e2e_probe2.py $P lag1 2>&1 | tail -1)"; done
  echo "$m fwd only P=8: $(python tools/e2e_probe2.py 8 2>&1 | grep 'lag1   grad=0')"
done
