#!/bin/bash
# bench.py with several BUILDS of the library (compile-time experiments: leaf pieces per curve, mailbox), one line each.
#   build here:   make -C vkhrt_b200/csrc OUT=../_lib/var_x EXTRA_NVFLAGS="-DVKHRT_LEAF_SPLIT_PHANTOM=4 ..."
#   on the box:   tools/lib_variants.sh tag "c2 c1" var_a var_b ...
# The variant library replaces the default one for the run and the default is put back afterwards.
tag=$1; wls=$2; shift 2
L=vkhrt_b200/_lib
mkdir -p gpurun_out
cp $L/libvkhrt_b200.so $L/libvkhrt_b200.so.default
for v in "$@"; do
  cp $L/$v/libvkhrt_b200.so $L/libvkhrt_b200.so
  for wl in $wls; do
    timeout 600 python bench.py --workload $wl --steps ${STEPS:-20} --warmup 5 --no-cpu-baseline --no-parity --no-strong-c5 > gpurun_out/${tag}_${v}_$wl.json 2> gpurun_out/${tag}_${v}_$wl.err
    python tools/variant_line.py "$v $wl" gpurun_out/${tag}_${v}_$wl.json
  done
done
mv $L/libvkhrt_b200.so.default $L/libvkhrt_b200.so
