#!/bin/bash
# Tuning variants of the headline instantiation group: recompiles csrc/_gen/inst_3.cu with extra -D flags and links
# it with the other objects of the current in-tree build into scratch/variants/libsvbasl_<name>.so
# (load with SVBASL_LIB=...).   usage: scratch/build_variant.sh <name> [-DFOO=1 ...]
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p scratch/variants /tmp/var_$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --ftz=true -Xcompiler -fPIC \
     --expt-relaxed-constexpr -I include "$@" -c svb_models_asl_b200/csrc/_gen/inst_${INST:-3}.cu -o /tmp/var_$name/inst_${INST:-3}.o
objs=$(ls svb_models_asl_b200/csrc/_obj/*.o | grep -v inst_${INST:-3}.o)
nvcc -shared -o scratch/variants/libsvbasl_$name.so $objs /tmp/var_$name/inst_${INST:-3}.o -lcudart
echo built scratch/variants/libsvbasl_$name.so
