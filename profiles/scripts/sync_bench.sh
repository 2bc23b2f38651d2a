python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python profiles/bench_sync.py > gpurun_out/sync.json 2> gpurun_out/sync.err; tail -5 gpurun_out/sync.err; python -c "
import json; d=json.load(open('gpurun_out/sync.json'))
print(d['rx_stream_host_capture_1shard'])
for r in d['rx_stream_device_capture']['by_shards']: print(r)
"
