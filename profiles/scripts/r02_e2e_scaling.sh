# round 2: name the e2e scaling limiter.  Run on an 8-GPU box: gpurun --gpus 8 -- 'bash profiles/scripts/r02_e2e_scaling.sh'
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | egrep "Model name|Socket|NUMA|^CPU\(s\)|Thread"; numactl -H 2>/dev/null | head -20; free -g | head -2; } > gpurun_out/r02_box_topology.txt 2>&1
: > gpurun_out/r02_pcie_probe_multi.jsonl
for n in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) profiles/scripts/pcie_probe_multi.py 2>/dev/null | tail -1 >> gpurun_out/r02_pcie_probe_multi.jsonl
done
cat gpurun_out/r02_pcie_probe_multi.jsonl
: > gpurun_out/r02_bench_scaling.jsonl
for n in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 3 --warmup 3 --no-cpu --frames 131072 2>/dev/null | tail -1 >> gpurun_out/r02_bench_scaling.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/r02_bench_scaling.jsonl"):
    d = json.loads(l); print(d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["tx_frac"])
PY
