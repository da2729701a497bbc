set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02d_gputest.log 2>&1; echo "pytest rc=$?" >> $O/r02d_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02d_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02d_smoke.log
DCB_BENCH_SHAPES=$O/r02d_per_shape_events.txt python bench.py --steps 20 --warmup 3 > $O/r02d_bench_line.json 2> $O/r02d_bench.err
python bench.py --steps 3 --warmup 3 --no-video --no-cpu-baseline --no-graph > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3100 -c 780 --csv --log-file $O/r02d_launches_one_step.csv python bench.py --steps 3 --warmup 3 --no-video --no-cpu-baseline --no-graph > $O/ncu_l.log 2>&1
python bench.py --config 1024 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02d_bench_1024.json 2> $O/r02d_bench_1024.err
tail -3 $O/r02d_gputest.log; tail -2 $O/r02d_smoke.log
