#!/bin/bash
# ncu of the build-side encoder (N3) on the stand-alone checker's timing cases (30 000 / 10 000 / 6 000 / 3 000 parents): the 84
# launches of the small verification cases are skipped.
mkdir -p gpurun_out tests/native/_build
EXE=tests/native/_build/neighbor_codes_gpu_check
make -C oracle port > /dev/null 2>&1
g++ -O1 -std=c++17 -I include -I oracle -I /usr/local/cuda/include tests/native/neighbor_codes_gpu_check.cpp \
    rabitq-ann-search_b200/cphnsw_b200/libcphnsw_b200.so oracle/libcphnsw_oracle.so -L/usr/local/cuda/lib64 -lcudart \
    -Wl,-rpath,$PWD/rabitq-ann-search_b200/cphnsw_b200 -Wl,-rpath,$PWD/oracle -o $EXE || exit 1
ncu --section SpeedOfLight --section Occupancy --section SchedulerStats --section WarpStateStats --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section LaunchStats \
    --clock-control none -k regex:neighbor_codes --launch-skip 84 -c 36 -f -o gpurun_out/n3_r02_big $EXE > gpurun_out/n3_ncu_big.log 2>&1
tail -2 gpurun_out/n3_ncu_big.log; ls -la gpurun_out/n3_r02_big.ncu-rep
