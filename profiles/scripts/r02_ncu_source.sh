# per-SASS-instruction executed counts / stall samples of the fft-512 kernels (source page of one ncu --set full capture)
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --frames 32768 --no-cpu --e2e-frames 2048 --oracle-frames 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rx_acquire512w|rx_demod512|tx512w" -s 3 -c 3 -o gpurun_out/src_main -f $B > gpurun_out/ncu_src_main.log 2>&1
tail -2 gpurun_out/ncu_src_main.log
ncu -i gpurun_out/src_main.ncu-rep --page source --csv > gpurun_out/src_main.csv 2>/dev/null
python profiles/scripts/ncu_brief.py gpurun_out/src_main.ncu-rep 32768 > gpurun_out/main_ncu_brief.txt
rm -f gpurun_out/*.ncu-rep
