#!/bin/bash
# compute-sanitizer is closed on the GPU pool, so memory safety of the persistent kernels is checked by the kernels themselves:
# builds libptb200_checked.so (-DPTB_CHECKED: device asserts on ring slots, chunk ranges, the wavefront kernel's pool / queues /
# stacks, tree node and sphere references; csrc/pt_device.cuh) and runs the whole GPU test-suite against it.  A failed assert
# prints file:line, traps, and fails the test.  Run on a GPU box:  bash tools/checked_tests.sh   (build here first: it travels)
set -e
cd "$(dirname "$0")/.."
python -c "from ascendpathtracing_b200 import build as b; print(b.build_variant('checked', ['PTB_CHECKED']))"
# second build: the fused resolve's rarely taken path -- a chunk claim postponed behind a straggling path, the ring running dry,
# idle lanes re-armed -- forced on every other claim (-DPTB_TEST_POSTPONE), asserts on; the fused images must not change
python -c "from ascendpathtracing_b200 import build as b; print(b.build_variant('postpone', ['PTB_TEST_POSTPONE', 'PTB_CHECKED']))"
if python -c "import torch, sys; sys.exit(0 if torch.cuda.is_available() else 1)"; then
    PTB200_LIB=$PWD/ascendpathtracing_b200/libptb200_checked.so python -m pytest tests -m gpu -q "$@"
    PTB200_LIB=$PWD/ascendpathtracing_b200/libptb200_postpone.so python -m pytest tests/test_gpu_fused_resolve.py -m gpu -q
fi
