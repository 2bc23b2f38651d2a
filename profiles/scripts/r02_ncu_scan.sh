ncu --set full --clock-control none --import-source on -k regex:stream_scan -s 2 -c 1 -o gpurun_out/prof_scan -f python profiles/bench_stream_dist.py --log2-samples 28 > gpurun_out/ncu_scan.log 2>&1
tail -2 gpurun_out/ncu_scan.log
