#!/bin/bash
# tools/ntt_variants.sh — rebuild the library with each NTT twiddle/occupancy variant and time the commit bench.
for v in "-DNTT2_MONT_TW=0 -DNTT2_MINBLOCKS=2" "-DNTT2_MONT_TW=1 -DNTT2_MINBLOCKS=2" "-DNTT2_MONT_TW=1 -DNTT2_MINBLOCKS=3" "-DNTT2_MONT_TW=1 -DNTT2_MINBLOCKS=4"; do
  BFGPU_NVCC_EXTRA="$v" python -c "import importlib; b=importlib.import_module('zkvm-brainfuck_b200.build'); b.build(force=True)" > /dev/null 2>&1
  echo "== $v"
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-prove 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['phases_ms_per_step'].items()}, d['root'][:2])"
done
