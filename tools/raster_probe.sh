#!/bin/bash
for g in 2 4 8 16 32; do echo "== raster $g"; LASER_B200_RASTER=$g timeout 200 python tools/packed_probe.py 2>&1 | grep -E "default row,row|packed A|x1"; done
