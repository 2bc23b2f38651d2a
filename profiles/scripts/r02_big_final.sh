# round 2, fft-4096 workload: final records after the layout-specialised instances (tests, smoke, bench line, batch sweep,
# kernel times by batch size, ncu summary + launch list of the three kernels)
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --workload big --steps 5 --warmup 3 > $O/r02_bench_big_1gpu.json 2> $O/r02_bench_big.err; tail -1 $O/r02_bench_big.err
: > $O/r02_bench_big_sweep.jsonl
for f in 1024 4096 65536; do python bench.py --workload big --frames $f --e2e-frames 1024 --steps 5 --warmup 3 --no-cpu --oracle-frames 0 >> $O/r02_bench_big_sweep.jsonl 2>> $O/r02_bench_big.err; done
BIG_BATCHES=64,256,1024,4096,16384 python profiles/bench_generic.py > $O/r02_big_path.json 2>> $O/r02_bench_big.err
BIG_BATCHES=4096 ncu --set full --clock-control none --import-source on -k regex:"big_demod|big_acquire|big_tx" -s 6 -c 3 -o $O/prof_big -f python profiles/bench_generic.py > $O/ncu_big.log 2>&1
python profiles/scripts/ncu_brief.py $O/prof_big.ncu-rep 4096 > $O/r02_big_ncu_summary.txt; rm -f $O/prof_big.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"big_demod|big_acquire|big_tx" -c 40 --csv --log-file $O/r02_launch_list_big.csv python bench.py --workload big --steps 2 --warmup 1 --frames 4096 --no-cpu --e2e-frames 512 --oracle-frames 0 > $O/ncu_big_ll.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_big_1gpu.json").read().strip().splitlines()[-1])
rf = d["roofline"]
print("big", round(d["value"]), d.get("rx_ms"), d.get("tx_ms"), rf.get("frac"), rf.get("tx_frac"), d.get("bit_errors"), d.get("oracle_check"), d["e2e"]["value"], d.get("clocks"), d.get("cpu_baseline"))
for l in open("gpurun_out/r02_bench_big_sweep.jsonl"):
    d = json.loads(l); print("big sweep", d["config"]["frames_per_gpu"], round(d["value"]), d["rx_ms"], d["roofline"]["frac"], d["roofline"]["tx_frac"], d["bit_errors"])
PY
grep "gpu__time\|per frame\|^==" $O/r02_big_ncu_summary.txt
