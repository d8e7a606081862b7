run() { python bench.py --steps 4 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); s=d['stages_ms']; print('$*', round(d['ms_per_step'],3), {k: s[k] for k in s if 'untimed' not in k})"; }
for l in 4 8 16; do run --lanes $l; done
for l in 4 8 16 32; do run --genes 60000 --samples 500 --design factorial --lanes $l; done
for l in 2 8 16 32; do run --genes 125000 --samples 1000 --lanes $l; done
