# A/B of tuning variants built with pydeseq2_b200/build.py (defines=..., out=...): bash scripts/variants_sweep.sh lib1.so lib2.so ...
run() { PDQ_LIB=pydeseq2_b200/$1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline "${@:2}" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); s=d['stages_ms']; print('$*', round(d['ms_per_step'],4), {k: s[k] for k in s if 'untimed' not in k})"; }
for lib in "$@"; do run $lib; done
for lib in "$@"; do run $lib --genes 60000 --samples 500 --design factorial; done
