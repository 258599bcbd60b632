#!/bin/bash
# A/B of the HP1 build variants on one box: python tools/build_variant.py <name> ... first, then
#   bash tools/run_variants.sh gpurun_out/variants.jsonl name1 name2 ...
out=$1; shift
: > $out
python tools/hp1_time.py >> $out 2>> $out.err
for n in "$@"; do
  AGX_LIB_PATH=tools/dbg/libagx_$n.so timeout 300 python tools/hp1_time.py >> $out 2>> $out.err
done
cat $out
