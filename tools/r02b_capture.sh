set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
DCB_BENCH_SHAPES=$O/r02b_per_shape_events.txt python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02b_bench_line.json 2> $O/r02b_bench.err
python bench.py --steps 3 --warmup 3 --no-video --no-cpu-baseline --no-graph > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3200 -c 800 --csv --log-file $O/r02b_launches_one_step.csv python bench.py --steps 3 --warmup 3 --no-video --no-cpu-baseline --no-graph > $O/ncu_l.log 2>&1
for w in cross:attn_tc temporal:attn_tc vae:vae_attn gn_apply:gn_stream_apply gemm_res:tap_gemm gemm_conv:tap_gemm; do
  n=${w%%:*}; k=${w##*:}
  python tools/prof_r02.py $n > $O/plain_$n.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $O/r02b_$n python tools/prof_r02.py $n > $O/ncu_$n.log 2>&1
done
python tools/bench_vae.py > $O/r02b_vae.txt 2>&1
python tools/bench_gn_stats.py > $O/r02b_gn_stats.txt 2>&1
ls -la $O
