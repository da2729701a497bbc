# same-box A/B of the step time: bash tools/ab_step.sh "ENV_A=.." "ENV_B=.." ...   (each arg = one variant's environment)
cd $GRAFT_REPO_ROOT
i=0
for v in "$@"; do
  i=$((i+1))
  env $v python bench.py --steps 20 --warmup 3 --no-video --no-cpu-baseline > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  python - "$v" gpurun_out/ab_$i.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); print(f"{sys.argv[1]:50s} ms/step {d['ms_per_step']:.3f}  e2e {1e3/d['e2e']['value']:.3f} ms  gemm {d['breakdown_ms'].get('tcgen05_gemm_conv')}  clk {d['clocks']['sm_mhz']} launches {d['launches_per_step']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
