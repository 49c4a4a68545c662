#!/bin/bash
# Round-2 FINAL ncu evidence for the training step (bench.py headline: yolov4, 8 images, 640x640), eager launches.
#   1) launch list of two steps (durations + DRAM bytes of every kernel, incl. torch's fills / copies)
#        -> launches_train_r02_final.csv   (the second half = one warm step)
#   2) --set full captures of the kernels that changed late in the round, exported as raw CSV pages
# Run on the GPU box:  bash profiles/capture_r02b.sh   (results land in gpurun_out/, summaries are copied to profiles/r02/)
export B2Y_NO_GRAPH=1
mkdir -p gpurun_out
CMD="python tools/train_step.py --steps 2"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/launches_train_r02_final.csv $CMD > gpurun_out/ncu_r02f_list.log 2>&1
wc -l gpurun_out/launches_train_r02_final.csv
full() {   # name regex skip count
    ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/r02f_$1 $CMD \
        > gpurun_out/ncu_r02f_$1.log 2>&1
    ncu -i gpurun_out/r02f_$1.ncu-rep --page raw --csv > gpurun_out/r02f_$1_raw.csv 2>/dev/null
    rm -f gpurun_out/r02f_$1.ncu-rep
    wc -l gpurun_out/r02f_$1_raw.csv
}
full wgrad_taps "wgrad_taps_kernel"         5 5      # all five narrow 3x3 weight gradients of step 2
full conv_fwd   "conv_tc_kernel"          222 8      # first forward convs of step 2 (narrow / HBM-bound layers, stats epilogue)
full bn_reduce  "bn_train_bwd_reduce"     107 4      # last layers of the net = first of the backward pass (small maps)
full maxpool    "maxpool_plane"             6 6      # SPP pooling forward + backward of step 2
