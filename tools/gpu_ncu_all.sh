#!/bin/bash
# One gpurun call: ncu --set full of one steady-state instance of every kernel on the path (plain run first, as required).
mkdir -p gpurun_out
set -x
python tools/glgm06_probe.py 200 0 1 > gpurun_out/plain_a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 52 -c 12 -f -o gpurun_out/prof_all_glgm06 python tools/glgm06_probe.py 200 0 1 > gpurun_out/ncu_a.log 2>&1
python tools/box_probe.py C3 400 > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 90 -c 6 -f -o gpurun_out/prof_all_box python tools/box_probe.py C3 400 > gpurun_out/ncu_b.log 2>&1
python tools/box_probe.py C3 400 > gpurun_out/plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"scale_copy|col_abs|axpby|box_matvec|sih_kernel|omega0" -c 8 -f -o gpurun_out/prof_all_disc python tools/box_probe.py C3 400 > gpurun_out/ncu_c.log 2>&1
python tools/ensemble_probe.py 512 400 > gpurun_out/plain_d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:homog_small -c 1 -f -o gpurun_out/prof_all_ens python tools/ensemble_probe.py 512 400 > gpurun_out/ncu_d.log 2>&1
tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log gpurun_out/ncu_d.log
