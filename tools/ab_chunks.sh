#!/bin/bash
# A/B of the persistent kernels' chunk size (paths per claim) on one B200: builds variants of libptb200 with other -D values
# here or on the GPU box (nvcc needed) and runs bench.py plus a quarter-size C4/C5 suite against each (PTB200_LIB selects).
# Result of round 1: profiles/r1_chunk_size_ab.md.
set -e
cd "$(dirname "$0")/.."
python - <<'PY'
from ascendpathtracing_b200 import build
build.build()
for name, defs in [("c64", ["PTB_CHUNK_BATCHES=64", "PTB_BVH_MAX_CHUNK=2048"]), ("c16", ["PTB_CHUNK_BATCHES=16", "PTB_BVH_MAX_CHUNK=256"]),
                   ("c4", ["PTB_CHUNK_BATCHES=4", "PTB_BVH_MAX_CHUNK=32"])]:
    build.build_variant(name, defs)
PY
for v in "" _c64 _c16 _c4 ""; do
    echo "variant ${v:-default (8 batches = 256 paths; BVH 64)}"
    export PTB200_LIB=$PWD/ascendpathtracing_b200/libptb200$v.so
    python bench.py --steps 30 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['pipeline']['ms_per_step'])"
    python tools/suite_multi_gpu.py --only c4,c5,c5mat --scale 0.25 --out gpurun_out/s.json 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['seconds'],4), round(d['mpaths_s']), round(d['grays_s'],2))"
done
