# fft-4096 demod: persistent clusters A/B (COFDM_BIG_PERSISTENT=0/1), GPU tests of the fft-4096 path first
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "generic_path or syncless or big" 2>&1 | tail -2
for p in 0 1; do COFDM_BIG_PERSISTENT=$p python bench.py --workload big --no-cpu --oracle-frames 64 --e2e-frames 256 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/big_per$p.json; done
python - <<'PY'
import json
for p in (0,1):
    d=json.load(open(f"gpurun_out/big_per{p}.json")); print("persistent",p, round(d["value"]), d["rx_ms"], d["tx_ms"], d["roofline"]["frac"], d["bit_errors"], d["oracle_check"]["frames_bytes_equal"], d["clocks"])
PY
