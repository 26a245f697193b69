#!/bin/bash
# generic: run the python command lines in $1.. each appended to gpurun_out/cmd.txt
mkdir -p gpurun_out; rm -f gpurun_out/cmd.txt
while [ $# -gt 0 ]; do
  echo "## $1" >> gpurun_out/cmd.txt
  timeout 300 bash -c "$1" >> gpurun_out/cmd.txt 2>&1
  shift
done
cat gpurun_out/cmd.txt
