# A/B of experimental builds (COFDM_LIB_PATH), same box, same run, interleaved
for rep in 1 2 3; do
for v in "" exp/old.so; do
  echo "variant: ${v:-product}"
  COFDM_LIB_PATH=${v:+$PWD/c-ofdm_b200/$v} python bench.py --steps 5 --warmup 3 --no-cpu --e2e-frames 8192 --oracle-frames 0 | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['rx_ms'], d['tx_ms'], d['roofline']['frac'], d['bit_errors'], d['clocks'])"
done
done
