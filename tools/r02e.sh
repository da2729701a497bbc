cd $GRAFT_REPO_ROOT
O=gpurun_out
DCB_GEMM_RAGGED256=0 DCB_GEMM_ORDER=0 python tools/bench_shapes.py > $O/r02e_a.txt 2>&1
DCB_GEMM_RAGGED256=0 DCB_GEMM_ORDER=1 python tools/bench_shapes.py > $O/r02e_b.txt 2>&1
DCB_GEMM_RAGGED_MIN=64 DCB_GEMM_ORDER=0 python tools/bench_shapes.py > $O/r02e_c.txt 2>&1
DCB_GEMM_RAGGED_MIN=64 DCB_GEMM_ORDER=1 python tools/bench_shapes.py > $O/r02e_d.txt 2>&1
python - <<'PY'
import re
cols=[]
for t in "abcd":
    cols.append([l.rstrip() for l in open(f"gpurun_out/r02e_{t}.txt")])
print("columns: a = old tiles, N-fastest | b = old tiles, wave-major | c = ragged256 (min 64), N-fastest | d = ragged256, wave-major   [us]")
for rows in zip(*cols):
    name=rows[0].split(":")[0]
    us=[re.search(r":\s+([\d.]+) us",r).group(1) for r in rows]
    print(f"{name:48s} " + "  ".join(f"{u:>7s}" for u in us))
PY
