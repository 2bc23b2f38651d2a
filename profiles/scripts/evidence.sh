# the round's evidence in one gpurun call (1 GPU): bench, launch list, full ncu captures of the three hot kernels, reference arm
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
B="python bench.py --steps 2 --warmup 1 --frames 32768 --no-cpu --e2e-frames 2048"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rx_acquire|rx_fused512|tx512" --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rx_acquire|rx_fused512|tx512" -s 3 -c 3 -o gpurun_out/prof_all -f $B > gpurun_out/ncu_all.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1
tail -c 400 gpurun_out/bench_final.json
