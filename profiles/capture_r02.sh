#!/bin/bash
# Round-2 ncu evidence for the training step (bench.py headline: yolov4, 8 images, 640x640), eager launches.
#   1) launch list of step 2 (durations + DRAM bytes of every kernel)           -> launches_train_r02.csv
#   2) --set full captures of a few launches per kernel family, exported as raw CSV pages (tensor-pipe %, DRAM GB/s ...)
# Run on the GPU box:  bash profiles/capture_r02.sh     (results land in gpurun_out/, summaries are copied to profiles/r02/)
export B2Y_NO_GRAPH=1
mkdir -p gpurun_out
CMD="python tools/train_step.py --steps 2"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 760 -c 760 --csv \
    --log-file gpurun_out/launches_train_r02.csv $CMD > gpurun_out/ncu_r02_list.log 2>&1
wc -l gpurun_out/launches_train_r02.csv
full() {   # name regex skip count
    ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/r02_$1 $CMD \
        > gpurun_out/ncu_r02_$1.log 2>&1
    ncu -i gpurun_out/r02_$1.ncu-rep --page raw --csv > gpurun_out/r02_$1_raw.csv 2>/dev/null
    rm -f gpurun_out/r02_$1.ncu-rep
    wc -l gpurun_out/r02_$1_raw.csv
}
full conv_fwd  "conv_tc_kernel"            274 6      # forward convs of step 2 (mid network)
full dgrad     "conv_tc_kernel"            383 6      # data gradients of step 2
full wgrad     "wgrad_tc_kernel"           139 6
full bn_fwd    "bn_train_fwd_kernel"       127 3
full bn_reduce "bn_train_bwd_reduce"       137 3
full bn_apply  "bn_train_bwd_apply"        137 3
