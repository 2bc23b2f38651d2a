# round 2, first GPU call: parity suite, old vs new demod kernel, ncu capture of the new kernel
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in 0 1; do echo "COFDM_RX_WARP=$w"; COFDM_RX_WARP=$w python bench.py --steps 5 --warmup 3 --no-cpu --e2e-frames 8192 | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['rx_ms'], d['tx_ms'], d['roofline']['frac'], d['roofline']['tx_frac'], d['bit_errors'])"; done
B="python bench.py --steps 2 --warmup 1 --frames 32768 --no-cpu --e2e-frames 2048"
ncu --set full --clock-control none --import-source on -k regex:"rx_demod512" -s 2 -c 1 -o gpurun_out/prof_dm1 -f $B > gpurun_out/ncu_dm1.log 2>&1
tail -2 gpurun_out/ncu_dm1.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rx_acquire|rx_demod512|rx_fused512|tx512" --csv --log-file gpurun_out/launches_r02a.csv $B > gpurun_out/ncu_ll.log 2>&1
grep -c rx_demod512 gpurun_out/launches_r02a.csv
