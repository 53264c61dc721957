#!/bin/bash
mkdir -p gpurun_out
python -m clive2_b200.movie render --scene bench-terrain-10m --width 1920 --height 1080 --samples 256 --seed 1 --out gpurun_out/r02r_c4_256spp.png > gpurun_out/r02r_render_c4.log 2>&1; echo "c4 rc=$?"; tail -1 gpurun_out/r02r_render_c4.log
python -m clive2_b200.movie render --scene bench-spheres --width 512 --height 512 --samples 256 --seed 1 --out gpurun_out/r02r_c2_256spp.png > gpurun_out/r02r_render_c2.log 2>&1; echo "c2 rc=$?"; tail -1 gpurun_out/r02r_render_c2.log
