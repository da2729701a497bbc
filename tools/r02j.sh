cd $GRAFT_REPO_ROOT
O=gpurun_out
python tools/bench_shapes.py > $O/r02j_a.txt 2>&1
DCB_GEMM_BN=192 python tools/bench_shapes.py > $O/r02j_b.txt 2>&1
DCB_GEMM_BN=128 python tools/bench_shapes.py > $O/r02j_c.txt 2>&1
python - <<'PY'
import re
cols=[[l.rstrip() for l in open(f"gpurun_out/r02j_{t}.txt")] for t in "abc"]
print("columns: default | BN=192 | BN=128 [us]")
for rows in zip(*cols):
    name=rows[0].split(":")[0]
    us=[re.search(r":\s+([\d.]+) us",r).group(1) for r in rows]
    print(f"{name:48s} " + "  ".join(f"{u:>7s}" for u in us))
PY
