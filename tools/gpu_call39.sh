#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_raster.py -q -x 2>&1 | tail -3 > gpurun_out/pytest_39.log
timeout 300 python tools/op_sweep.py --raster-only --sizes 256,512 --out gpurun_out/raster_sweep.json > gpurun_out/raster_sweep.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"gray|image_to|agent" --log-file gpurun_out/raster_launches.csv python tools/op_sweep.py --raster-only --sizes 256 > /dev/null 2>&1
