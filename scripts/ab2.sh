#!/bin/bash
# scripts/ab2.sh workload rounds lib1 lib2 ... : A/B timing of library variants ("default" = the in-tree build)
wl=$1; rounds=$2; shift; shift
for i in $(seq $rounds); do
  for lib in "$@"; do
    if [ "$lib" = default ]; then unset TORJ_CUDA_LIB; else export TORJ_CUDA_LIB=$lib; fi
    timeout 300 python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline $BENCH_EXTRA 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl $lib', round(d['ms_per_step'],2), d['roofline']['counters']['n_acc'], d['rays_ok'])"
  done
done
