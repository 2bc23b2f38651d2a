set -x
mkdir -p gpurun_out
BIG_BATCHES=4096 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"big_tx" -s 2 -c 1 -o gpurun_out/src_bigtx -f python profiles/bench_generic.py > gpurun_out/ncu_src_bigtx.log 2>&1
tail -2 gpurun_out/ncu_src_bigtx.log
ncu -i gpurun_out/src_bigtx.ncu-rep --page source --csv > gpurun_out/src_bigtx.csv 2>/dev/null
python profiles/scripts/ncu_brief.py gpurun_out/src_bigtx.ncu-rep 4096 > gpurun_out/bigtx_ncu_brief.txt
rm -f gpurun_out/*.ncu-rep
