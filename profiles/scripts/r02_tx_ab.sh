# round 2: tx kernel A/B (COFDM_TX_WARP=1: one warp per symbol, tx512w.cuh; 0: the two-warp-team kernel), same box, same run
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tx or align or facade or config" 2>&1 | tail -3
for v in 1 0 1 0; do
  echo "COFDM_TX_WARP=$v"
  COFDM_TX_WARP=$v python bench.py --steps 5 --warmup 3 --no-cpu --e2e-frames 8192 | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['rx_ms'], d['tx_ms'], d['roofline']['frac'], d['roofline']['tx_frac'], d['bit_errors'], d['e2e']['value'], d['clocks'])"
done
