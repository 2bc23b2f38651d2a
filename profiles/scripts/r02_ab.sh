# A/B of experimental builds of the demod kernel (COFDM_LIB_PATH), same box, same run
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_and_device" 2>&1 | grep -E "Error|error|assert" | head -8
for v in "" exp/direct.so exp/direct3.so exp/minb3.so; do
  echo "variant: ${v:-product}"
  for i in 1 2; do
  COFDM_LIB_PATH=${v:+$PWD/c-ofdm_b200/$v} python bench.py --steps 5 --warmup 3 --no-cpu --e2e-frames 8192 | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['rx_ms'], d['tx_ms'], d['roofline']['frac'], d['bit_errors'])"
  done
done
