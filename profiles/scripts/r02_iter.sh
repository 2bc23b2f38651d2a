# round 2 iteration: parity suite, bench, ncu capture of the demod kernel
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 5 --warmup 3 --no-cpu --e2e-frames 8192 | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['rx_ms'], d['tx_ms'], d['roofline']['frac'], d['roofline']['tx_frac'], d['bit_errors'])"; done
B="python bench.py --steps 2 --warmup 1 --frames 32768 --no-cpu --e2e-frames 2048"
ncu --set full --clock-control none --import-source on -k regex:"rx_demod512|rx_acquire" -s 2 -c 2 -o gpurun_out/prof_dm -f $B > gpurun_out/ncu_dm.log 2>&1
tail -1 gpurun_out/ncu_dm.log
