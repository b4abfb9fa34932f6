#!/bin/bash
# round-2 ncu evidence (run under gpurun on one B200); outputs under gpurun_out/
set -x
O=gpurun_out
# 1. every launch of the default tile workload (f16x3, 16 tiles/step), eager launches, no extras
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file $O/r2_launches_tile_f16x3_b16.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > $O/cap1.log 2>&1
# 2. dram traffic of the conv launches of one step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -s 219 -c 73 --csv \
    --log-file $O/r2_conv_tc_traffic_f16x3_b16.csv python tools/trace_tc.py 16 f16x3 > $O/cap2.log 2>&1
# 3. full-set captures: head tower conv (f16x3), deformable conv (f16x3), 1x1 + residual (f16x3)
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 5 -c 1 -o $O/r2_tower_f16x3 \
    python tools/one_conv.py 256 256 3 1 1 128 128 16 0 f16x3 > $O/cap3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 5 -c 1 -o $O/r2_dcn_f16x3 \
    python tools/one_conv.py 256 256 3 1 1 128 128 16 0 f16x3 1 > $O/cap4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 5 -c 1 -o $O/r2_res1x1_f16x3 \
    python tools/one_conv.py 64 256 1 1 0 256 256 16 1 f16x3 > $O/cap5.log 2>&1
# 4. NMS: launch list + full set of the sweep and the lazy resolve, dense 100k
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2_launches_nms100k_dense.csv \
    python bench.py --workload nms_100k --steps 2 --warmup 3 --no-cpu-baseline > $O/cap6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nms_sweep_kernel -s 3 -c 1 -o $O/r2_nms_sweep_dense100k \
    python bench.py --workload nms_100k --steps 2 --warmup 3 --no-cpu-baseline > $O/cap7.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nms_resolve_lazy -s 3 -c 1 -o $O/r2_nms_resolve_dense100k \
    python bench.py --workload nms_100k --steps 2 --warmup 3 --no-cpu-baseline > $O/cap8.log 2>&1
python tools/trace_tc.py 1 f16x3 > $O/trace_f16x3_b1.log 2>&1
ORP_TC_TRACE=1 python tools/trace_tc.py 1 f16x3 2>&1 | grep "^tc\[" > $O/trace_f16x3_b1_launches.log
tail -n 3 $O/trace_f16x3_b1.log
# 5. Swin-T (f16x3, 8 tiles): every launch of one dense step
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_swin_f16x3_b8.csv python tools/trace_tc.py 8 f16x3 swin_tiny > $O/cap9.log 2>&1
ORP_TC_TRACE=1 python tools/trace_tc.py 16 f16x3 2>&1 | grep "^tc\[" > $O/trace_f16x3_b16_launches.log
