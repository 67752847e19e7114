set -u
O=gpurun_out
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2j_bench_plain.json 2> $O/r2j_bench_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2j_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2j_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 300 python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-e2e > $O/r2j_bench_eager.json 2> $O/r2j_bench_eager.err && \
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k "regex:(hashgrid_fwd|hashgrid_bwd|mlp_fwd_tc|mlp_bwd_tc|composite_fwd|composite_bwd_sweep|lpf_loss|compact_kernel|march_kernel)" \
  --launch-skip 20 -c 12 -o $O/r2j_full -f python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-e2e > $O/r2j_ncu_full.log 2>&1; echo "full capture rc=$?"
du -sh $O
