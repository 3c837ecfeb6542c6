#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"grid_remap_tiles|grid_masks|grid_tile_lists|grid_cells" -s 8 -c 4 \
    -o gpurun_out/r02_full -f python bench.py --steps 1 --warmup 1 --batch 32 --kernel-only > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out/r02_full*
