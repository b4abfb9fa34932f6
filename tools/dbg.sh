A="256 64 1 1 0 256 256 16 0 f16x3"
run() { echo "--- $*"; env "$@" python tools/one_conv.py $A 2>&1 | tail -1 | cut -c1-150; }
run ORP_TC_NO_NCAT=1
run ORP_TC_NO_NCAT=1 ORP_TC_NO_PDL=1
run ORP_TC_NO_NCAT=1 ORP_TC_NO_MERGE=1
run ORP_TC_NO_NCAT=1 ORP_TC_NO_BRES=1
run ORP_TC_NO_NCAT=1 ORP_TC_STAGES=2
run X=1
run ORP_TC_NO_PDL=1
A="64 256 1 1 0 256 256 16 0 f16x3"
run X=1
run ORP_TC_NO_PDL=1
python tools/trace_tc.py 16 f16x3 2>&1 | tail -2 | cut -c1-200
ORP_TC_NO_PDL=1 python tools/trace_tc.py 16 f16x3 2>&1 | tail -2 | cut -c1-200
