# fft-4096 path: GPU tests, bench of the big workload, ncu summary of its kernels
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
COFDM_BIG_LAY=0 python bench.py --workload big --no-cpu 2>/dev/null | tail -1 > gpurun_out/big_lay0.json
python bench.py --workload big --no-cpu 2>/dev/null | tail -1 > gpurun_out/big_lay1.json
python - <<'PY'
import json
for f in ("big_lay0","big_lay1"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["value"], d["roofline"]["frac"], d["roofline"]["tx_frac"], d.get("rx_ms"), d.get("tx_ms"), d.get("bit_errors"), d.get("oracle_check"), d["clocks"])
PY
BIG_BATCHES=4096 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"big_tx" -s 2 -c 1 -o gpurun_out/src_bigtx -f python profiles/bench_generic.py > gpurun_out/ncu_src_bigtx.log 2>&1
tail -2 gpurun_out/ncu_src_bigtx.log
ncu -i gpurun_out/src_bigtx.ncu-rep --page source --csv > gpurun_out/src_bigtx2.csv 2>/dev/null
python profiles/scripts/ncu_brief.py gpurun_out/src_bigtx.ncu-rep 4096 > gpurun_out/bigtx_ncu_brief.txt
rm -f gpurun_out/*.ncu-rep
