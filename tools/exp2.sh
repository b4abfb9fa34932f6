for d in 0 1 2 4; do echo dbg $d; ORP_TC_DCN_DEBUG=$d python tools/one_conv.py 256 256 3 1 1 128 128 16 0 f16x3 1; done
