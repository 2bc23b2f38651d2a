B="python bench.py --steps 1 --warmup 1 --frames 32768 --no-cpu --e2e-frames 2048"
$B > gpurun_out/pre.json 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"rx_acquire|rx_fused512" -s 2 -c 2 -o gpurun_out/prof_acq2 -f $B > gpurun_out/ncu_acq.log 2>&1
tail -3 gpurun_out/ncu_acq.log
