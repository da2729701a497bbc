cd $GRAFT_REPO_ROOT
N=${NPROC:-2}
run() { TAG="$1" env $1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/shard_bisect.py 2>&1 | grep "^SHARD\|Error\|error" | head -3; }
run "DCB_X=default"
run "DCB_FUSED_GN_STATS=0"
run "DCB_SHARD_V1=1"
run "POISON=0"
