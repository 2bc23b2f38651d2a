BIG_BATCHES=4096 ncu --set full --clock-control none --import-source on -k regex:"big_demod|big_acquire" -s 8 -c 2 -o gpurun_out/prof_big -f python profiles/bench_generic.py > gpurun_out/ncu_big.log 2>&1
tail -2 gpurun_out/ncu_big.log
