for mb in 3 4; do
ORP_NMS_SWEEP_MINB=$mb python bench.py --workload nms_100k --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('minb', $mb, d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['kept'], d['config']['candidates'])"
done
for mb in 3 4 6; do
ORP_NMS_SWEEP_MINB=$mb python tools/trace_tc.py 16 f16x3 2>&1 | grep -i "whole step\|post"
done
