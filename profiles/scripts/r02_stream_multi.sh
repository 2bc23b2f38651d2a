# round 2: BASELINE configs[3] on 1/2/4/8 GPUs, one 2^30-sample capture (strong scaling).  gpurun --gpus 8 -- 'bash profiles/scripts/r02_stream_multi.sh'
: > gpurun_out/r02_stream_multi_gpu.jsonl
for n in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) profiles/bench_stream_dist.py 2> gpurun_out/stream_$n.err | tail -1 >> gpurun_out/r02_stream_multi_gpu.jsonl
  tail -2 gpurun_out/stream_$n.err
done
cat gpurun_out/r02_stream_multi_gpu.jsonl
# BASELINE configs[4] on 2/4/8 GPUs (bench.py --workload big; the 1-GPU line is profiles/r02_bench_big_1gpu.json)
: > gpurun_out/r02_bench_big_multi.jsonl
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + n)) bench.py --workload big --gpus $n --steps 5 --warmup 3 --no-cpu 2> gpurun_out/big_$n.err | tail -1 >> gpurun_out/r02_bench_big_multi.jsonl
  tail -1 gpurun_out/big_$n.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r02_bench_big_multi.jsonl"):
    d = json.loads(l); print(d["n_gpus"], round(d["value"]), d["roofline"]["frac"], d["roofline"]["tx_frac"], d["bit_errors"], round(d["e2e"]["value"]))
PY
