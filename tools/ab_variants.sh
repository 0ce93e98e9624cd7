# usage: bash tools/ab_variants.sh name1 name2 ...   (variants built by tools/build_variants.py)
# runs the hot-loop timings of every variant and the sliding parity tests of the last one
for v in "$@"; do
  export CUSTMA_LIB=$PWD/custereomatching_b200/_variants/libcustma_$v.so
  for ph in ${PHASES:-bwd}; do
    for w in kitti cfg3 cfg2; do
      echo "== $v $ph $w: $(timeout 300 python tools/run_hot.py --phase $ph --workload $w --iters 20 2>&1 | tail -1)"
    done
  done
done
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
