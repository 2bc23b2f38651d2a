B="python bench.py --steps 2 --warmup 1 --frames 32768 --no-cpu --e2e-frames 2048"
ncu --set full --clock-control none --import-source on -k regex:"rx_acquire|rx_fused512|tx512" -s 3 -c 3 -o gpurun_out/prof_all -f $B > gpurun_out/ncu_all.log 2>&1
tail -1 gpurun_out/ncu_all.log
