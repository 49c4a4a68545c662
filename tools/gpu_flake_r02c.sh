#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/flake_probe.log
: > $L
run() { echo "=== $*" >> $L; ( env "$@" timeout 150 python tools/flake_probe.py 5 >> $L 2>&1 ); echo "rc=$?" >> $L; }
run B2Y_X=default
run B2Y_PDL=0
run B2Y_WGRAD_STREAM=0
run B2Y_EPI_TMA=0
run B2Y_GRAD_FIRSTWRITE=0
echo "=== graph mode" >> $L; timeout 150 python tools/flake_probe.py 4 graph >> $L 2>&1
grep -v "Model Summary" $L
