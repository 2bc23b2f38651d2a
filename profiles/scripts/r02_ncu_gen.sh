# the any-size kernels (generic.cuh) on the fft-4096 / 64-QAM configuration (COFDM_BIG=0 takes it off the cluster kernels):
# one ncu --set full capture of each kernel at a 512-frame batch (= one sub-batch of launch_rx_generic)
mkdir -p gpurun_out
COFDM_BIG=0 BIG_BATCHES=512 timeout 300 ncu --set full --clock-control none -k regex:"gen_" -s 5 -c 5 -o gpurun_out/prof_gen -f python profiles/bench_generic.py > gpurun_out/ncu_gen.log 2>&1
tail -2 gpurun_out/ncu_gen.log
python profiles/scripts/ncu_brief.py gpurun_out/prof_gen.ncu-rep 512 > gpurun_out/r02_gen_ncu_summary.txt
rm -f gpurun_out/*.ncu-rep
grep "^==\|gpu__time\|per frame" gpurun_out/r02_gen_ncu_summary.txt
