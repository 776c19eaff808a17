#!/bin/bash
mkdir -p gpurun_out
python juzhen_b200/cpp/build_dropin.py --extract-datasets
cd build/dropin/project
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 3000 -c 1000 --csv \
   --log-file ../../../gpurun_out/mnist_launches_z.csv ../bin/demo_mnist > /dev/null 2>&1
wc -l ../../../gpurun_out/mnist_launches_z.csv
