# round 2 evidence: ncu --set full of the three fft-512 hot kernels, the fft-4096 kernels and the synchronisation kernels.
# The reports are summarised on the GPU box (profiles/scripts/ncu_brief.py) and deleted: gpurun pulls at most 64 MiB back.
B="python bench.py --steps 2 --warmup 1 --frames 32768 --no-cpu --e2e-frames 2048 --oracle-frames 0"
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"rx_acquire512w|rx_demod512|tx512w" -s 3 -c 3 -o $O/prof_r02_main -f $B > $O/ncu_r02_main.log 2>&1
python profiles/scripts/ncu_brief.py $O/prof_r02_main.ncu-rep 32768 > $O/r02_final_ncu_summary.txt
BIG_BATCHES=4096 ncu --set full --clock-control none -k regex:"big_demod|big_acquire|big_tx" -s 12 -c 3 -o $O/prof_r02_big -f python profiles/bench_generic.py > $O/ncu_r02_big.log 2>&1
python profiles/scripts/ncu_brief.py $O/prof_r02_big.ncu-rep 4096 > $O/r02_big_ncu_summary.txt; rm -f $O/prof_r02_big.ncu-rep
: > $O/r02_sync_ncu_summary.txt
for k in t2sin_metric2 preamble_corr4 stream_scan stream_gather; do
  ncu --set full --clock-control none -k regex:$k -s 1 -c 1 -o $O/prof_r02_$k -f python profiles/bench_sync.py > $O/ncu_r02_$k.log 2>&1
  python profiles/scripts/ncu_brief.py $O/prof_r02_$k.ncu-rep 1 >> $O/r02_sync_ncu_summary.txt; rm -f $O/prof_r02_$k.ncu-rep
done
# launch list of the default bench (per-launch durations; shares of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tx512w|rx_acquire512w|rx_demod512" -c 60 --csv --log-file $O/r02_launch_list.csv $B > $O/ncu_r02_ll.log 2>&1
ls -la $O | head -30
