#!/bin/bash
# Full GPU suite, then ncu of the build-side encoder (N3) over the stand-alone checker's cases (sections, not --set full: 40 launches).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02b.log
EXE=tests/native/_build/neighbor_codes_gpu_check
if [ -x $EXE ]; then
  $EXE > gpurun_out/n3_check_r02.log 2>&1; tail -3 gpurun_out/n3_check_r02.log
  ncu --section SpeedOfLight --section Occupancy --section SchedulerStats --section WarpStateStats --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section LaunchStats \
      --clock-control none -k regex:neighbor_codes -c 80 -f -o gpurun_out/n3_r02 $EXE > gpurun_out/n3_ncu.log 2>&1
  ls -la gpurun_out/n3_r02.ncu-rep
fi
