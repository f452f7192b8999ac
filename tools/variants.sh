#!/bin/bash
# build variant libraries here (CPU box): tools/variants.sh build tag1:"-DX=1" tag2:"-DX=2" ...  -> vkhrt_b200/_lib/variants/<tag>.so
# run on the GPU box:                     tools/variants.sh run "<bench args>" tag1 tag2 ...       (restores the default library)
mode=$1; shift
V=vkhrt_b200/_lib/variants
if [ $mode = build ]; then
  mkdir -p $V
  for spec in "$@"; do
    tag=${spec%%:*}; flags=${spec#*:}
    make -s -C vkhrt_b200/csrc OUT=../_lib/var_$tag EXTRA_NVFLAGS="$flags" ../_lib/var_$tag/libvkhrt_b200.so && cp vkhrt_b200/_lib/var_$tag/libvkhrt_b200.so $V/$tag.so && rm -rf vkhrt_b200/_lib/var_$tag
    echo "built $tag ($flags)"
  done
else
  args=$1; shift
  cp vkhrt_b200/_lib/libvkhrt_b200.so /tmp/default.so
  for tag in "$@"; do
    [ $tag = default ] && cp /tmp/default.so vkhrt_b200/_lib/libvkhrt_b200.so || cp $V/$tag.so vkhrt_b200/_lib/libvkhrt_b200.so
    for rep in 1 2; do
      python bench.py $args --no-cpu-baseline --no-parity --no-strong-c5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', '%.1f Mrays/s  e2e %.1f' % (d['value'], d['e2e']['value']))"
    done
  done
  cp /tmp/default.so vkhrt_b200/_lib/libvkhrt_b200.so
fi
