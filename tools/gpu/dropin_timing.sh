#!/bin/bash
# The honest drop-in number: the reference's own UNMODIFIED drivers built against the shim (integration/build_dropin.sh),
# timed by their own time_it (wall_clock_timer.h:82-95), next to the reference CPU binaries on the same host.
mkdir -p gpurun_out
D=oracle/_ref/dropin
{
  echo "# rs-b200 . 19 4096 (unmodified RS.cpp: two MFA_NTT calls through the shim + its CPU scaling loop, pageable host memory)"
  $D/rs-b200 . 19 4096 2>&1 | tail -4
  echo "# the same with FASTECC_B200_NO_PIN=1 (array left pageable: driver-staged copies)"
  FASTECC_B200_NO_PIN=1 $D/rs-b200 . 19 4096 2>&1 | tail -2
  echo "# ntt-b200 n 19 4096"
  $D/ntt-b200 n 19 4096 2>&1 | tail -3
  echo "# reference CPU build (AVX2 + OpenMP, all host threads): rs-avx2 . 19 4096"
  OMP_WAIT_POLICY=active oracle/_ref/rs-avx2 . 19 4096 2>&1 | tail -3
} > gpurun_out/dropin_timing.log 2>&1
cat gpurun_out/dropin_timing.log
