#!/bin/bash
# Round-2 profile capture (run on the GPU box through gpurun; ONE ncu-using call per part).  Every ncu pass repeats a
# command that has just exited 0 without ncu; numbers printed under ncu are never used as bench values.
# usage: profile_round2.sh yolo | conv | membound
set -u
O=gpurun_out/prof2
mkdir -p $O
case "${1:-yolo}" in
yolo)
  # launch list + DRAM traffic of one step of the benchmark (default arithmetic), per mode
  for m in bf16x3 tf32; do
    python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary --math $m > $O/bench_short_$m.json 2> $O/bench_short_$m.err || exit 1
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 648 -c 220 --csv \
        --log-file $O/launches_$m.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary --math $m > $O/ncu_launch_$m.log 2>&1
  done
  ;;
conv)
  # full captures of one launch of the channels-innermost kernel: CTA pairs, 256 -> 512 3x3 @26^2, batch 64
  for m in bf16x3 tf32 3xtf32; do
    TNN_NHWC=1 python tools/conv_one.py $m 64,256,26,512,3,1 > $O/one_$m.log 2>&1 && \
    TNN_NHWC=1 ncu --set full --clock-control none --import-source on -k regex:k_conv_nhwc -c 1 -s 3 -o $O/r02_nhwc_pair_c256f512_$m \
        python tools/conv_one.py $m 64,256,26,512,3,1 > $O/ncu_$m.log 2>&1
  done
  ;;
membound)
  python tools/membound_driver.py > $O/membound_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'k_rows|k_map_vec|k_channel_reduce|k_plane|k_sgd|k_upsample_nhwc|k_nchw_to_nhwc' \
      -s 9 -c 9 -o $O/r02_membound python tools/membound_driver.py > $O/ncu_membound.log 2>&1
  ;;
esac
ls -la $O
