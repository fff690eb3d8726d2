#!/bin/bash
# One `ncu --set full` capture of one kernel at launch skip+1 of tools/ncu_target.py (default: the growth-phase wave of ~1.1
# million records that profiles/r01e/ used), reduced on the GPU box to details / raw metrics / SASS-source summary / per-CUDA-line
# summary.   bash tools/ncu_one.sh <tag> <name> <kernel regex> [skip]        (environment knobs such as PB_SAMPLE_G pass through)
tag=$1; name=$2; regex=$3; skip=${4:-20}
out=gpurun_out/ncu_$tag
mkdir -p $out /tmp/ncu
PB_GRAPH=0 timeout -k 10 400 ncu --set full --import-source on --clock-control none -k regex:$regex -s $skip -c 1 -o /tmp/ncu/$name -f python tools/ncu_target.py > $out/$name.log 2>&1
ncu -i /tmp/ncu/$name.ncu-rep --page details > $out/${name}_details.txt 2>/dev/null
ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > $out/${name}_raw.csv 2>/dev/null
ncu -i /tmp/ncu/$name.ncu-rep --page source --csv --print-source sass 2>/dev/null | python tools/ncu_source_summary.py > $out/${name}_source_summary.txt
ncu -i /tmp/ncu/$name.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null > $out/${name}_cuda_sass.csv
python tools/ncu_line_summary.py < $out/${name}_cuda_sass.csv > $out/${name}_line_summary.txt 2>&1
gzip -f $out/${name}_raw.csv $out/${name}_cuda_sass.csv
echo "$name: $(grep -m1 'Duration' $out/${name}_details.txt | tr -s ' ')  $(grep -m1 'dram__bytes_read.sum ' $out/${name}_details.txt | tr -s ' ')"
