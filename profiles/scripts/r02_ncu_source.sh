# per-SASS-instruction executed counts / stall samples of the two demod kernels (source page of one ncu --set full capture each)
set -x
mkdir -p gpurun_out
BIG_BATCHES=4096 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"big_demod" -s 4 -c 1 -o gpurun_out/src_big -f python profiles/bench_generic.py > gpurun_out/ncu_src_big.log 2>&1
tail -2 gpurun_out/ncu_src_big.log
ncu -i gpurun_out/src_big.ncu-rep --page source --csv > gpurun_out/src_big_demod.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rx_demod512" -s 4 -c 1 -o gpurun_out/src_rx -f python bench.py --frames 32768 --steps 2 --warmup 1 --no-cpu --oracle-frames 0 > gpurun_out/ncu_src_rx.log 2>&1
tail -2 gpurun_out/ncu_src_rx.log
ncu -i gpurun_out/src_rx.ncu-rep --page source --csv > gpurun_out/src_rx_demod.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -8
