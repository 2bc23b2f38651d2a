# round 2: the fft-4096 path (big.cuh): parity, batch sweep, A/B against the any-size acquisition (COFDM_BIG_ACQUIRE=0) and the old path (COFDM_BIG=0)
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generic" 2>&1 | tail -3
show='import json,sys; d=json.loads(sys.stdin.read())
for r in d["by_batch"]: print(r["frames"], "rx_ms", round(r["rx_ms"],3), "rx_frac", round(r["rx_frac"],3), "acq", round(r.get("rx_acquire_ms",0),3), "dem", round(r.get("rx_demod_ms",0),3), "tx_ms", round(r["tx_ms"],3), "tx_frac", round(r["tx_frac"],3), "bad", r["frames_with_errors"])'
python profiles/bench_generic.py | tee gpurun_out/r02_big_path.json | python -c "$show"
echo "COFDM_BIG_ACQUIRE=0"; COFDM_BIG_ACQUIRE=0 python profiles/bench_generic.py | python -c "$show"
