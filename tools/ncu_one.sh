mkdir -p gpurun_out/ncu_r01f /tmp/ncu
name=k_sample_stream
timeout -k 10 400 ncu --set full --import-source on --clock-control none -k regex:k_sample_stream -s 20 -c 1 -o /tmp/ncu/$name -f python tools/ncu_target.py > gpurun_out/ncu_r01f/$name.log 2>&1
ncu -i /tmp/ncu/$name.ncu-rep --page details > gpurun_out/ncu_r01f/${name}_details.txt 2>/dev/null
ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu_r01f/${name}_raw.csv 2>/dev/null
ncu -i /tmp/ncu/$name.ncu-rep --page source --csv --print-source sass 2>/dev/null | python tools/ncu_source_summary.py > gpurun_out/ncu_r01f/${name}_source_summary.txt
gzip -f gpurun_out/ncu_r01f/${name}_raw.csv
grep -m3 "Duration\|Executed Ipc\|Registers" gpurun_out/ncu_r01f/${name}_details.txt
head -50 gpurun_out/ncu_r01f/${name}_source_summary.txt
