#!/bin/bash
# compute-sanitizer is closed on the GPU pool, so memory safety of the persistent kernels is checked by the kernels themselves:
# builds libptb200_checked.so (-DPTB_CHECKED: device asserts on ring slots, chunk ranges, the wavefront kernel's pool / queues /
# stacks, tree node and sphere references; csrc/pt_device.cuh) and runs the whole GPU test-suite against it.  A failed assert
# prints file:line, traps, and fails the test.  Run on a GPU box:  bash tools/checked_tests.sh   (build here first: it travels)
set -e
cd "$(dirname "$0")/.."
python -c "from ascendpathtracing_b200 import build as b; print(b.build_variant('checked', ['PTB_CHECKED']))"
if python -c "import torch, sys; sys.exit(0 if torch.cuda.is_available() else 1)"; then
    PTB200_LIB=$PWD/ascendpathtracing_b200/libptb200_checked.so python -m pytest tests -m gpu -q "$@"
fi
