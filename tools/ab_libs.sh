#!/bin/bash
# A/B of libptb200 build variants (ascendpathtracing_b200/build.py: build_variant) on the C2 trace kernel:
#   tools/ab_libs.sh out.txt name1 name2 ...     ("base" = the shipped library)
# Each variant: the bit-exact parity tests (-k trace), then bench.py's kernel time twice, alternating.
out=$1; shift
: > "$out"
for name in "$@"; do
  lib=ascendpathtracing_b200/libptb200_$name.so; [ "$name" = base ] && lib=ascendpathtracing_b200/libptb200.so
  PTB200_LIB=$PWD/$lib python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "trace or c2_full or open_random or special" 2>&1 | tail -1 | sed "s/^/$name parity: /" >> "$out"
done
for rep in 1 2; do
  for name in "$@"; do
    lib=ascendpathtracing_b200/libptb200_$name.so; [ "$name" = base ] && lib=ascendpathtracing_b200/libptb200.so
    PTB200_LIB=$PWD/$lib python bench.py --steps 100 --warmup 5 --no-e2e --no-strong --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$name rep $rep: kernel_ms %.4f  ms_per_step %.4f  frac %.4f' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))" >> "$out"
  done
done
cat "$out"
