#!/bin/bash
# Builds the standalone timing tools against the kernel sources (sm_100a only).  usage: tools/build_tools.sh [name] [-D...]
# e.g. tools/build_tools.sh rec_test ; tools/build_tools.sh rec_test_base -DMVAE_SV_BULK=0 -DMVAE_SV_HP=1
set -e
cd "$(dirname "$0")/.."
NAME=${1:-rec_test}; shift || true
C=molecular-vae_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo "$@" \
  tools/rec_test.cu $C/gru_rec2.cu $C/gru_rec.cu $C/errors.cu -o tools/$NAME
echo built tools/$NAME
