#!/bin/bash
# Round-2 confirmation on an 8 x B200 box (gpurun --gpus 8): topology, the multi-device tests on real devices, bench.py at
# N = 1 and N = 8 (strong C3 job in both), and the C++ host binary on BASELINE configs C3 / C4 / C5 / C2 over 8 GPUs.
# Everything lands in gpurun_out/r2_n8_*.
O=gpurun_out
N=${1:-8}
{ nvidia-smi topo -m; nvidia-smi --query-gpu=index,pci.bus_id,name --format=csv; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; cat /sys/devices/system/node/node*/cpulist 2>/dev/null; free -g | head -2; } > $O/r2_n8_topology.txt 2>&1
python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r2_n8_tests.log 2>&1; echo "tests rc=$?" >> $O/r2_n8_tests.log; tail -2 $O/r2_n8_tests.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/r2_n8_bench_n1.json 2> $O/r2_n8_bench_n1.err; echo "bench n1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > $O/r2_n8_bench_n$N.json 2> $O/r2_n8_bench_n$N.err; echo "bench n$N rc=$?"
W=$(mktemp -d); mkdir -p $W/input $W/output; EXE=$PWD/ascendpathtracing_b200/render_gpu; OUT=$PWD/$O
( cd $W
  for g in 1 $N; do
    $EXE --image --gen --counter-rng 1 --width 3840 --height 2160 --samples 256 --gpus $g --reps 3 --p6 --json $OUT/r2_n8_host_c3_g$g.json | tail -1
    $EXE --image --gen --scene-kind random:10000 --materials --bvh --gamma --counter-rng 1 --width 1920 --height 1080 --samples 64 --gpus $g --reps 3 --p6 --json $OUT/r2_n8_host_c4_g$g.json | tail -1
    $EXE --image --gen --counter-rng 1 --width 1920 --height 1080 --samples 128 --depth 50 --gpus $g --reps 3 --p6 --json $OUT/r2_n8_host_c5d50_g$g.json | tail -1
    $EXE --gen --counter-rng 1 --width 1024 --height 768 --samples 16 --gpus $g --reps 3 --json $OUT/r2_n8_host_c2_dropin_g$g.json | tail -1
    md5sum output/color.bin | sed "s/^/c2 dropin g$g color.bin /"
  done )
rm -rf $W
