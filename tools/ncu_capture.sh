#!/bin/bash
# Round evidence: one `ncu --set full` capture per kernel at the plateau wave of BASELINE config 2 (and of the dark pass of
# config 3), reduced ON THE GPU BOX to text (details page, raw metrics, per-opcode / per-stall source aggregation) so that
# the results fit gpurun's 64 MiB return limit.  Usage (from the repo root, under gpurun):  bash tools/ncu_capture.sh r01e
tag=${1:-r01}
out=gpurun_out/ncu_$tag
mkdir -p $out /tmp/ncu
cap() {   # name, regex, skip, extra-ncu-args..., -- target args
  name=$1; regex=$2; skip=$3; shift 3
  ncu --set full --import-source on --clock-control none -k regex:$regex -s $skip -c 1 -o /tmp/ncu/$name -f "$@" > $out/$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page details > $out/${name}_details.txt 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > $out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page source --csv --print-source sass 2>/dev/null | python tools/ncu_source_summary.py > $out/${name}_source_summary.txt
  echo "$name: $(grep -m1 -A0 'Duration' $out/${name}_details.txt | tr -s ' ')"
}
for k in k_loop k_sample k_emit k_finalize k_bucket_fill k_bucket_scan; do cap $k $k 20 python tools/ncu_target.py; done
for k in k_dark_prepare k_sample k_dark_emit; do cap dark_$k $k 0 --profile-from-start off python tools/ncu_target.py --dark; done
cp /tmp/ncu/k_loop.ncu-rep /tmp/ncu/k_sample.ncu-rep $out/ 2>/dev/null
du -sh $out
