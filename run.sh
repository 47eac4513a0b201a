#!/bin/bash
# run.sh -- same command line as the reference's run.sh (reference run.sh:32-60):
#   bash run.sh -r|--run-mode {gpu,cpu,sim,npu} [-v|--soc-version <SOC>] [-i|--install-path <path>]
# plus the problem size, which the reference fixes at compile time (src/common.h:4-6):
#   [-W width] [-H height] [-S samples] [-D depth]
# and, gpu mode only, what the BASELINE configs beyond the reference's own need:
#   [-g|--gpus N]        N GPUs of this box (drop-in files: the reference's blockDim-way slice split; --image: strided columns + P2P gather)
#   [-I|--image]         production path: input/spheres.bin -> output/color.ppm, no rays.bin / color.bin (4K x 1024 spp would be 204 GB)
#   [-K|--scene-kind K]  default | smallpt | random:N[:SEED]   (what is written to input/spheres.bin; see include/ptb200.h: ptb200_scene_layout)
#   [-M|--materials] [-B|--bvh] [-G|--gamma] [-C|--counter-rng SEED] [-R|--reps N] [-X|--max-depth N] [--p6]
#   e.g. C3:  bash run.sh -r gpu -I -W 3840 -H 2160 -S 256 -g 8 -C 1 --p6
#        C4:  bash run.sh -r gpu -I -K random:10000 -M -B -G -W 1920 -H 1080 -S 64 -g 8 -C 1 --p6
#        C5:  bash run.sh -r gpu -I -W 1920 -H 1080 -S 128 -D 50 -g 8 -C 1 --p6
#
#   gpu  B200 build: libptb200.so + render_gpu; rays generated on the device (bit-identical replay of
#        scripts/gen_data.py's seed-0 stream), kernel, device resolve -> output/color.bin, output/color.ppm
#   cpu  the reference's own cpu mode for comparison: its src/*.cpp compiled against oracle/shim
#        (needs /root/reference or a prebuilt oracle/_ref artefact); inputs/outputs through the oracle tools
#   sim, npu  Ascend-only modes of the reference: refused here
CURRENT_DIR=$(cd "$(dirname "${BASH_SOURCE:-$0}")" && pwd)
cd "$CURRENT_DIR" || exit 1

RUN_MODE=gpu
SOC_VERSION=B200
WIDTH=16; HEIGHT=16; SAMPLES=1; DEPTH=5
GPUS=1; EXTRA=()
OPTS=$(getopt -a --options r:v:i:W:H:S:D:g:IK:MBGC:R:X: --longoptions run-mode:,soc-version:,install-path:,width:,height:,samples:,depth:,gpus:,image,scene-kind:,materials,bvh,gamma,counter-rng:,reps:,max-depth:,p6 -- "$@") || exit 1
eval set -- "$OPTS"
while :; do
    case "$1" in
    -r | --run-mode) RUN_MODE="$2"; shift 2 ;;
    -v | --soc-version) SOC_VERSION="$2"; shift 2 ;;
    -i | --install-path) shift 2 ;;   # CANN install path: accepted, unused
    -W | --width) WIDTH="$2"; shift 2 ;;
    -H | --height) HEIGHT="$2"; shift 2 ;;
    -S | --samples) SAMPLES="$2"; shift 2 ;;
    -D | --depth) DEPTH="$2"; shift 2 ;;
    -g | --gpus) GPUS="$2"; shift 2 ;;
    -I | --image) EXTRA+=(--image); shift ;;
    -K | --scene-kind) EXTRA+=(--scene-kind "$2"); shift 2 ;;
    -M | --materials) EXTRA+=(--materials); shift ;;
    -B | --bvh) EXTRA+=(--bvh); shift ;;
    -G | --gamma) EXTRA+=(--gamma); shift ;;
    -C | --counter-rng) EXTRA+=(--counter-rng "$2"); shift 2 ;;
    -R | --reps) EXTRA+=(--reps "$2"); shift 2 ;;
    -X | --max-depth) EXTRA+=(--max-depth "$2"); shift 2 ;;
    --p6) EXTRA+=(--p6); shift ;;
    --) shift; break ;;
    *) echo "[ERROR] Unexpected option: $1"; exit 1 ;;
    esac
done

RUN_MODE_LIST="gpu cpu sim npu"
if [[ " $RUN_MODE_LIST " != *" $RUN_MODE "* ]]; then
    echo "ERROR: RUN_MODE error, This build supports gpu (B200) and cpu (reference cpu mode)!"
    exit 1
fi
if [ "$RUN_MODE" = "sim" ] || [ "$RUN_MODE" = "npu" ]; then
    echo "ERROR: run mode '$RUN_MODE' needs the Ascend toolchain; this is the B200 build (use -r gpu or -r cpu)"
    exit 1
fi

set -e
mkdir -p input output
rm -rf input/*.bin output/*.bin output/*.ppm output/report.json

if [ "$RUN_MODE" = "gpu" ]; then
    python3 -c "import __graft_entry__ as g; from ascendpathtracing_b200 import build as b; from ascendpathtracing_b200.host import build as h; b.build(); h.build()"
    echo "INFO: compile op on ${RUN_MODE} succeed! (soc ${SOC_VERSION})"
    ./ascendpathtracing_b200/render_gpu --width "$WIDTH" --height "$HEIGHT" --samples "$SAMPLES" --depth "$DEPTH" --gpus "$GPUS" --gen --ppm \
        --json output/report.json "${EXTRA[@]}"
    echo "INFO: execute op on ${RUN_MODE} succeed!"
else
    python3 -m oracle.tools cpu-mode --width "$WIDTH" --height "$HEIGHT" --samples "$SAMPLES" --depth "$DEPTH"
    echo "INFO: execute op on ${RUN_MODE} succeed!"
fi
