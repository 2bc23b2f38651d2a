# fft-4096 demod: cluster scheduling policy preference A/B (driver default / spread / load balancing = the library's choice), then
# the bench line of the big workload with the library's default and the GPU tests of the fft-4096 path
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "generic_path or syncless or big" 2>&1 | tail -2
for p in 0 1 2; do COFDM_BIG_CLUSTER_POLICY=$p python bench.py --workload big --no-cpu --oracle-frames 0 --e2e-frames 256 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/big_pol$p.json; done
python bench.py --workload big --steps 5 --warmup 3 > gpurun_out/r02_bench_big_1gpu.json 2> gpurun_out/r02_bench_big.err
python - <<'PY'
import json
for p in (0,1,2):
    d=json.load(open(f"gpurun_out/big_pol{p}.json")); print("policy",p, round(d["value"]), d["rx_ms"], d["tx_ms"], d["roofline"]["frac"], d["bit_errors"], d["clocks"])
d = json.loads(open("gpurun_out/r02_bench_big_1gpu.json").read().strip().splitlines()[-1])
rf = d["roofline"]
print("big", round(d["value"]), d.get("rx_ms"), d.get("tx_ms"), rf.get("frac"), rf.get("tx_frac"), rf.get("achieved"), rf.get("tx_kernel_gbs"), d.get("bit_errors"), d.get("oracle_check"), d["e2e"]["value"], d.get("clocks"))
PY
