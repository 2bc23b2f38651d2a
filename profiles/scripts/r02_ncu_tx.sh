B="python bench.py --steps 2 --warmup 1 --frames 32768 --no-cpu --e2e-frames 2048"
ncu --set full --clock-control none --import-source on -k regex:"tx512w" -s 2 -c 1 -o gpurun_out/prof_txw -f $B > gpurun_out/ncu_txw.log 2>&1
tail -2 gpurun_out/ncu_txw.log
