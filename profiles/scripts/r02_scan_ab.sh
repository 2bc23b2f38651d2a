# stream scanner: shards per GPU and resident CTAs per SM (COFDM_SCAN_MINB builds), one 2^30-sample capture on one GPU
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or apps" 2>&1 | tail -2
show='import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d["frames_after_merge"], d["payload_ok_all_ranks_incl_overlap"], d["unmerged_boundaries"], "ms", round(d["seconds"]*1e3,2), "scan_ms", d["scan_kernel_ms_max_rank"], d["stage_ms_rank0"])'
for v in ":444" "exp/w5b6.so:888" "exp/w5b6.so:1776" "exp/w5b8.so:1184"; do
  lib=${v%%:*}; sh=${v##*:}
  echo "variant ${lib:-product} shards $sh"
  COFDM_LIB_PATH=${lib:+$PWD/c-ofdm_b200/$lib} python profiles/bench_stream_dist.py --shards $sh 2>/dev/null | python -c "$show"
done
