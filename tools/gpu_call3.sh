#!/bin/bash
mkdir -p gpurun_out/r2
python tools/ncu_target.py conv dw > gpurun_out/r2/ncu_target_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"halo_conv_tc_kernel|dw_tile_kernel" --launch-skip 4 -c 4 \
    -o gpurun_out/r2/r2_hot -f python tools/ncu_target.py conv dw > gpurun_out/r2/ncu_hot.log 2>&1
tail -5 gpurun_out/r2/ncu_hot.log; ls -la gpurun_out/r2/
