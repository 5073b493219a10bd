#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> : plain run, launch list, full capture of the three tc16 kernels of one warm step
tag=${1:-r02x}
mkdir -p gpurun_out
python tools/step_once.py 5 > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 6 --csv --log-file gpurun_out/${tag}_launches.csv python tools/step_once.py 5 > gpurun_out/${tag}_ncu1.log 2>&1
python tools/step_once.py 5 > gpurun_out/${tag}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 9 -c 3 -o gpurun_out/${tag}_prof -f python tools/step_once.py 5 > gpurun_out/${tag}_ncu2.log 2>&1
tail -3 gpurun_out/${tag}_ncu2.log
ls -la gpurun_out/${tag}_prof.ncu-rep
