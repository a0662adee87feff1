#!/bin/bash
# Development build of a kernel variant (Panda only): tools/build_variant.sh <name> [extra nvcc -D flags...]
# -> build_variants/libtrk_<name>.so ; run with TRK_B200_LIB=build_variants/libtrk_<name>.so
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/build_variants; mkdir -p $out/obj_$name
cs=$root/torch_robotics_b200/csrc
common="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DTRK_ONLY_PANDA $@"
nvcc $common -c -o $out/obj_$name/trk_b200.o $cs/trk_b200.cu &
nvcc $common -DTRK_INST_ROBOT=robots::Panda -c -o $out/obj_$name/panda.o $cs/trk_static_inst.cu &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libtrk_$name.so $out/obj_$name/trk_b200.o $out/obj_$name/panda.o
echo built $out/libtrk_$name.so
