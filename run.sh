#!/bin/bash
# run.sh -- same command line as the reference's run.sh (reference run.sh:32-60):
#   bash run.sh -r|--run-mode {gpu,cpu,sim,npu} [-v|--soc-version <SOC>] [-i|--install-path <path>]
# plus the problem size, which the reference fixes at compile time (src/common.h:4-6):
#   [-W width] [-H height] [-S samples] [-D depth]
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
OPTS=$(getopt -a --options r:v:i:W:H:S:D: --longoptions run-mode:,soc-version:,install-path:,width:,height:,samples:,depth: -- "$@") || exit 1
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
rm -rf input/*.bin output/*.bin output/*.ppm

if [ "$RUN_MODE" = "gpu" ]; then
    python3 -c "import __graft_entry__ as g; from ascendpathtracing_b200 import build as b; from ascendpathtracing_b200.host import build as h; b.build(); h.build()"
    echo "INFO: compile op on ${RUN_MODE} succeed! (soc ${SOC_VERSION})"
    ./ascendpathtracing_b200/render_gpu --width "$WIDTH" --height "$HEIGHT" --samples "$SAMPLES" --depth "$DEPTH" --gen --ppm
    echo "INFO: execute op on ${RUN_MODE} succeed!"
else
    python3 -m oracle.tools cpu-mode --width "$WIDTH" --height "$HEIGHT" --samples "$SAMPLES" --depth "$DEPTH"
    echo "INFO: execute op on ${RUN_MODE} succeed!"
fi
