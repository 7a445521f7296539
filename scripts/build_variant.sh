#!/bin/bash
# Build a library variant for A/B runs: scripts/build_variant.sh NAME "-DFLAG=1 ..."  -> ab/NAME.so (tuning aid)
set -e
NAME=$1; FLAGS=$2
T=/tmp/var_$NAME
rm -rf $T; mkdir -p $T
cp -r Makefile include opencl-raytracing_b200 $T/
mkdir -p $T/oracle $T/tests/hostsim
( cd $T && make -j8 NVCCFLAGS_EXTRA="$FLAGS" ${FASTDIV+FASTDIV="$FASTDIV"} opencl-raytracing_b200/libraytracing_cuda.so >/dev/null 2>&1 )
mkdir -p ab
cp $T/opencl-raytracing_b200/libraytracing_cuda.so ab/$NAME.so
echo "ab/$NAME.so built with '$FLAGS'"
