#!/bin/bash
# Round-2 profile capture (run on the GPU box through gpurun; ONE ncu-using call per part).  Every ncu pass repeats a
# command that has just exited 0 without ncu; numbers printed under ncu are never used as bench values.
# usage: profile_round2.sh yolo | conv | split | planes | membound
set -u
O=gpurun_out/prof2
mkdir -p $O
case "${1:-yolo}" in
yolo)
  # launch list + DRAM traffic of one step of the benchmark (default arithmetic), per mode
  for m in bf16x3 tf32; do
    python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary --math $m > $O/bench_short_$m.json 2> $O/bench_short_$m.err || exit 1
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 500 -c 400 --csv \
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
split)
  # the benchmark's default path: CTA pairs on pre-split bf16 planes (no converter warps), 256 -> 512 3x3 @26^2, batch 64;
  # and the tensor-core image layer
  TNN_SPLIT=1 python tools/conv_one.py bf16x3 64,256,26,512,3,1 > $O/one_split.log 2>&1 && \
  TNN_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:k_conv_nhwc -c 1 -s 3 -o $O/r02_nhwc_pair_presplit_c256f512_bf16x3 \
      python tools/conv_one.py bf16x3 64,256,26,512,3,1 > $O/ncu_split.log 2>&1
  python tools/image_one.py bf16x3 planes > $O/one_image.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_conv_image -c 1 -s 2 -o $O/r02_image_c3f32_bf16x3 \
      python tools/image_one.py bf16x3 planes > $O/ncu_image.log 2>&1
  ;;
planes)
  # the benchmark's default graph since the second half of round 2: every tensor only as planes.  (1) a wide 3x3 layer with
  # the folded shortcut's operand read from planes, (2) a 1x1 layer on planes (no converter warps)
  TNN_SPLIT=2 python tools/conv_one.py bf16x3 64,256,26,512,3,1 > $O/one_planes_res.log 2>&1 && \
  TNN_SPLIT=2 ncu --set full --clock-control none --import-source on -k regex:k_conv_nhwc -c 1 -s 3 -o $O/r02_nhwc_pair_planes_res_c256f512_bf16x3 \
      python tools/conv_one.py bf16x3 64,256,26,512,3,1 > $O/ncu_planes_res.log 2>&1
  TNN_SPLIT=1 python tools/conv_one.py bf16x3 64,256,52,128,1,1 > $O/one_planes_1x1.log 2>&1 && \
  TNN_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:k_conv_nhwc -c 1 -s 3 -o $O/r02_nhwc_pair_planes_1x1_c256f128_bf16x3 \
      python tools/conv_one.py bf16x3 64,256,52,128,1,1 > $O/ncu_planes_1x1.log 2>&1
  ;;
membound)
  python tools/membound_driver.py > $O/membound_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'k_rows|k_map_vec|k_channel_reduce|k_plane|k_sgd|k_upsample_nhwc|k_nchw_to_nhwc' \
      -s 9 -c 9 -o $O/r02_membound python tools/membound_driver.py > $O/ncu_membound.log 2>&1
  ;;
esac
# the .ncu-rep files (source import) are tens of MB each and gpurun brings back at most 64 MiB: keep their raw / source
# pages as CSV and drop the reports
for r in $O/*.ncu-rep; do
  [ -f "$r" ] || continue
  ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null
  ncu -i $r --page source --csv > ${r%.ncu-rep}.source.csv 2>/dev/null
  rm -f $r
done
ls -la $O
