# single-GPU evidence of the final build -> gpurun_out/final/ (copied into profiles/r02_* afterwards)
set -x
O=gpurun_out/final; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/gputest.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/smoke.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/ref.err
python bench.py --steps 30 --warmup 5 > $O/r02_bench_kitti8.json 2> $O/kitti8.err
for w in cfg1 cfg2 cfg3 cfg5band verify15; do python bench.py --steps 30 --warmup 5 --workload $w --no-cpu-baseline > $O/r02_bench_$w.json 2> $O/$w.err; done
python bench.py --steps 30 --warmup 5 --workload cfg2 --graph --no-cpu-baseline > $O/r02_bench_cfg2_graph.json 2> $O/cfg2g.err
python bench.py --steps 30 --warmup 5 --graph --no-cpu-baseline --no-e2e > $O/r02_bench_kitti8_graph.json 2> $O/k8g.err
for d in shifted natural; do python bench.py --steps 30 --warmup 5 --data $d --no-cpu-baseline > $O/r02_bench_kitti8_$d.json 2> $O/$d.err; done
python bench.py --steps 10 --warmup 3 --scaling strong --global-pairs 64 --no-cpu-baseline --no-e2e > $O/r02_bench_strong_n1.json 2> $O/strong.err
python tools/run_head.py > $O/r02_head.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_fwd_bwd_kitti8.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sliding_forward_kernel|sliding_backward_kernel" -c 4 -o $O/r02_main_kernels python tools/run_hot.py --phase both --iters 1 > $O/ncu_full.log 2>&1
ls -la $O
