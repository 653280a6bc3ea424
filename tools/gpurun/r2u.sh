#!/bin/bash
# round 2, call U: LayerNorm backward with the residual operand prefetched vs the previous build (same box), blocks-per-SM knob
cd /root/repo
MVLT_LIB=/root/repo/mvlt_b200/lib/libmvlt_b200_oldnorm.so python tools/ln_sweep.py | sed 's/^/old  /'
python tools/ln_sweep.py | sed 's/^/new  /'
MVLT_LN_BWD_BPS=2 python tools/ln_sweep.py | sed 's/^/new  /'
MVLT_LN_BWD_BPS=3 python tools/ln_sweep.py | sed 's/^/new  /'
MVLT_LN_BWD_BPS=4 python tools/ln_sweep.py | sed 's/^/new  /'
MVLT_LN_BWD_PASSES=8 python tools/ln_sweep.py | sed 's/^/new  /'
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "layernorm or norm" 2>&1 | tail -2
