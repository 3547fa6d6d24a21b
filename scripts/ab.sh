#!/bin/bash
# A/B timing of library variants on the bench bundle: scripts/ab.sh rounds lib1 lib2 ...   ("default" = the in-tree build)
rounds=$1; shift
for i in $(seq $rounds); do
  for lib in "$@"; do
    if [ "$lib" = default ]; then unset TORJ_CUDA_LIB; else export TORJ_CUDA_LIB=$lib; fi
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['ms_per_step'],2), d['absorbed_fraction'], d['roofline']['counters']['n_acc'])"
  done
done
