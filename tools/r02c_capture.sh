set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02c_gputest.log 2>&1; echo "pytest rc=$?" >> $O/r02c_gputest.log
DCB_BENCH_SHAPES=$O/r02c_per_shape_events.txt python bench.py --steps 20 --warmup 3 > $O/r02c_bench_line.json 2> $O/r02c_bench.err
python bench.py --steps 3 --warmup 3 --no-video --no-cpu-baseline --no-graph > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3100 -c 780 --csv --log-file $O/r02c_launches_one_step.csv python bench.py --steps 3 --warmup 3 --no-video --no-cpu-baseline --no-graph > $O/ncu_l.log 2>&1
python tools/prof_gemm.py 5120 1280 1280 f32 res > $O/plain_g.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tap_gemm -s 3 -c 1 -f -o $O/r02c_gemm_bn192 python tools/prof_gemm.py 5120 1280 1280 f32 res > $O/ncu_g.log 2>&1
python tools/bench_shapes.py > $O/r02c_shapes.txt 2>&1
python tools/bench_vae.py > $O/r02c_vae.txt 2>&1
python tools/bench_attn.py > $O/r02c_attn.txt 2>&1
python bench.py --config 1024 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02c_bench_1024.json 2> $O/r02c_bench_1024.err
python bench.py --config 256 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02c_bench_256.json 2> $O/r02c_bench_256.err
tail -3 $O/r02c_gputest.log
