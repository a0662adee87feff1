mkdir -p gpurun_out
run() {
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 --skip-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', d['value'], d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'coll', d['collective']['collective_ms'], 'strong', d['strong']['value'], d['strong']['ms_per_step'])"
}
run default
run default2
NCCL_MAX_NCHANNELS=1 NCCL_MIN_NCHANNELS=1 run nch1
NCCL_MAX_CTAS=2 run maxctas2
TRK_BENCH_REDUCE=main run main
