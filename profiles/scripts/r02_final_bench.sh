# round 2 final records: tests, smoke, 1-GPU bench (default + QPSK + big + big batch sweep), reference arm, parity stress,
# sync kernels, ncu summary + launch list of the fft-512 kernels
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; tail -1 $O/r02_bench_1gpu.err
python bench.py --steps 5 --warmup 3 --mod-type 2 --no-cpu > $O/r02_bench_1gpu_qpsk.json 2>> $O/r02_bench_1gpu.err
python bench.py --workload big --steps 5 --warmup 3 > $O/r02_bench_big_1gpu.json 2>> $O/r02_bench_1gpu.err
: > $O/r02_bench_big_sweep.jsonl
for f in 1024 4096 65536; do python bench.py --workload big --frames $f --e2e-frames 1024 --steps 5 --warmup 3 --no-cpu --oracle-frames 0 >> $O/r02_bench_big_sweep.jsonl 2>> $O/r02_bench_1gpu.err; done
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2>> $O/r02_bench_1gpu.err
python profiles/scripts/stress_parity.py > $O/r02_parity_stress.json 2> $O/r02_parity_stress.err; tail -1 $O/r02_parity_stress.err
python profiles/bench_sync.py > $O/r02_sync_kernels.json 2>> $O/r02_bench_1gpu.err
BIG_BATCHES=64,256,1024,4096,16384 python profiles/bench_generic.py > $O/r02_big_path.json 2>> $O/r02_bench_1gpu.err
B="python bench.py --steps 2 --warmup 1 --frames 32768 --no-cpu --e2e-frames 2048 --oracle-frames 0"
ncu --set full --clock-control none --import-source on -k regex:"rx_acquire512w|rx_demod512|tx512w" -s 3 -c 3 -o $O/prof_r02_main -f $B > $O/ncu_r02_main.log 2>&1
python profiles/scripts/ncu_brief.py $O/prof_r02_main.ncu-rep 32768 > $O/r02_final_ncu_summary.txt; rm -f $O/prof_r02_main.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tx512w|rx_acquire512w|rx_demod512" -c 60 --csv --log-file $O/r02_launch_list.csv $B > $O/ncu_r02_ll.log 2>&1
python - <<'PY'
import json
for f in ("r02_bench_1gpu.json", "r02_bench_1gpu_qpsk.json", "r02_bench_big_1gpu.json", "r02_bench_reference_arm.json"):
    d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
    rf = d.get("roofline", {})
    print(f, round(d["value"]), d.get("rx_ms"), d.get("tx_ms"), rf.get("frac"), rf.get("tx_frac"), d.get("bit_errors"), d.get("oracle_check"), d["e2e"]["value"], d.get("rx_ms_with_ambiguity_count_on"), d.get("clocks"))
for l in open("gpurun_out/r02_bench_big_sweep.jsonl"):
    d = json.loads(l); print("big sweep", d["config"]["frames_per_gpu"], round(d["value"]), d["rx_ms"], d["roofline"]["frac"], d["roofline"]["tx_frac"], d["bit_errors"])
st = json.loads(open("gpurun_out/r02_parity_stress.json").read())
print("stress frames", sum(r["frames"] for r in st), "worst", max(max(r["synced"], r["grid"], r["constell"], r["chan"]) for r in st), "shift_mismatch", sum(r["shift_mismatch"] for r in st), "differing", sum(r["differing"] for r in st))
PY
