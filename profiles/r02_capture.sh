#!/bin/bash
# Round-2 evidence run (one gpurun call): tests, smoke, bench lines, ncu launch list, ncu --set full capture.
# Every ncu pass runs only after the same command has exited 0 without ncu.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/r2f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2f_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > $O/r2f_bench_default.json 2> $O/r2f_bench_default.err; echo "bench rc=$?"
timeout 200 python profiles/time_mlp.py 524288 > $O/r2f_time_mlp.log 2>&1; tail -3 $O/r2f_time_mlp.log
for w in synthetic_pb_off synthetic_budget render_sweep; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > $O/r2f_bench_$w.json 2> $O/r2f_bench_$w.err; echo "bench $w rc=$?"
done
timeout 600 python bench.py --emulate-ranks 8 --no-cpu-baseline > $O/r2f_bench_share8.json 2> $O/r2f_bench_share8.err; echo "share8 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2f_bench_ref.json 2> $O/r2f_bench_ref.err; echo "ref rc=$?"
# launch list of the bench command (plain run first)
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2f_bench_plain.json 2> $O/r2f_bench_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2f_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2f_ncu_launches.log 2>&1; echo "launch list rc=$?"
# (eds: run on its own, ~3 min: python bench.py --workload eds --no-cpu-baseline)
# full capture of the rated kernels; gpurun brings back at most 64 MiB: one step's worth of launches (eager step so that every launch is a kernel node ncu can replay)
timeout 300 python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-e2e > $O/r2f_bench_eager.json 2> $O/r2f_bench_eager.err && \
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k "regex:(hashgrid_fwd|hashgrid_bwd|mlp_fwd_tc|mlp_bwd_tc|composite_fwd|composite_bwd_sweep|lpf_loss|compact_kernel|march_kernel)" \
  --launch-skip 20 -c 12 -o $O/r2f_full -f python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-e2e > $O/r2f_ncu_full.log 2>&1; echo "full capture rc=$?"
