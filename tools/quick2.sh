timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for w in kitti cfg3 cfg2 cfg1; do echo "== bwd $w: $(timeout 300 python tools/run_hot.py --phase bwd --workload $w --iters 20 2>&1 | tail -1)"; done
