#!/bin/bash
# --set full capture of the tcgen05 halo attention kernels on the Halo-T* stage-1 shape (override with HT_ONLY=<Hs>)
mkdir -p gpurun_out
export HT_ONLY=${HT_ONLY:-56}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attn_ht_(fwd|bwd)" -s 13 -c 2 -o gpurun_out/prof_ht python tools/bench_haloattn.py > gpurun_out/ncu_ht.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/ncu_ht.log
