for rep in 1 2; do for d in 1 2; do for c in 1536 2048 3072 4096 8192; do
  echo -n "rep $rep depth $d chunk $c: "; COFDM_PIPE_DEPTH=$d COFDM_PIPE_CHUNK=$c python bench.py --steps 5 --warmup 2 --frames 65536 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['e2e']['value'], d['e2e']['ms_per_step'])"
done; done; done
