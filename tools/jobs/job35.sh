#!/bin/bash
# captures of the final build of round 2: launch lists (config 2 / 3), full capture of the remap and prep kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 22 -c 22 --csv \
    --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --batch 256 --kernel-only > gpurun_out/r2_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"grid_remap_tiles|grid_masks|grid_tile_lists|grid_cells" -s 8 -c 4 \
    -o gpurun_out/r02_full python bench.py --steps 1 --warmup 1 --batch 32 --kernel-only > gpurun_out/r2_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 24 -c 26 --csv \
    --log-file gpurun_out/r02_launches_chain.csv python bench.py --config 3 --steps 2 --warmup 1 --batch 256 --kernel-only > gpurun_out/r2_ncu_list3.log 2>&1
python tools/timeline_probe.py > gpurun_out/r2_timeline.log 2>&1
ls -la gpurun_out/r02_*
