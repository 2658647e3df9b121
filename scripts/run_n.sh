# usage: scripts/run_n.sh N [bench args...]   -> prints a compact summary line
N=$1; shift
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" 2> gpurun_out/run_n_$N.log | tee gpurun_out/run_n_$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=%d ms=%.2f value=%.3e frac=%.4f e2e_ms=%.1f launches=%d' % (d['n_gpus'], d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['gpu_launches']))"
grep "shard" gpurun_out/run_n_$N.log | head -8
