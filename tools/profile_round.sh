#!/bin/bash
# Round profile capture (run on the GPU box through gpurun).  Every ncu pass repeats a command that has just exited 0
# without ncu; numbers printed under ncu are never used as bench values.
# usage: profile_round.sh [yolo] [cifar] [sgemm]   (default: all three; gpurun brings back at most 64 MiB, and the five
# full captures together exceed that, so the parts can be run in separate calls)
set -u
O=gpurun_out/prof
mkdir -p $O
PARTS="${*:-yolo cifar sgemm}"
has() { case " $PARTS " in *" $1 "*) return 0;; *) return 1;; esac; }
if has yolo; then
python bench.py --steps 5 --warmup 3 --layers-out $O/conv_layers.json > $O/bench_line.json 2> $O/bench.err || exit 1
python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_short.json 2>/dev/null || exit 1
# (1) launch list: one full step
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_launch.log 2>&1
# (2) DRAM traffic of the convolution kernels of one step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'k_umma|k_conv_direct' -s 225 -c 75 --csv --log-file $O/conv_dram.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_dram.log 2>&1
# (3) full captures of one launch of each tensor-core kernel
python tools/conv_one.py 3xtf32 64,256,26,512,3,1 > $O/one_gather.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_umma -c 1 -s 2 -o $O/r01_gather_c256f512 \
    python tools/conv_one.py 3xtf32 64,256,26,512,3,1 > $O/ncu_gather.log 2>&1
python tools/conv_one.py 3xtf32 64,64,104,128,3,1 > $O/one_halo.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_umma_halo -c 1 -s 2 -o $O/r01_halo_c64f128 \
    python tools/conv_one.py 3xtf32 64,64,104,128,3,1 > $O/ncu_halo.log 2>&1
python tools/conv_one.py 3xtf32 64,256,52,128,1,1 > $O/one_tma.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_umma -c 1 -s 2 -o $O/r01_tma1x1_c256f128 \
    python tools/conv_one.py 3xtf32 64,256,52,128,1,1 > $O/ncu_tma.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu --math tf32 > $O/bench_tf32.json 2>/dev/null
fi
if has cifar; then
# (4) CIFAR-10 training
python tools/cifar_train.py --batch 128 --steps 50 > $O/cifar_b128.json 2>$O/cifar.err
python tools/cifar_train.py --batch 128 --steps 50 --no-graph > $O/cifar_b128_nograph.json 2>>$O/cifar.err
python tools/cifar_train.py --batch 512 --steps 30 > $O/cifar_b512.json 2>>$O/cifar.err
python tools/cifar_train.py --batch 128 --steps 30 --unfused > $O/cifar_b128_unfused.json 2>>$O/cifar.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/cifar_launches.csv \
    python tools/cifar_train.py --batch 128 --steps 4 --warmup 3 --no-graph > $O/ncu_cifar.log 2>&1
python tools/conv_one.py 3xtf32 128,32,32,32,3,1 > /dev/null 2>&1 && \
TNN_BWD=1 ncu --set full --clock-control none --import-source on -k regex:k_umma_wgrad -c 1 -s 1 -o $O/r01_wgrad_c32f32 \
    python tools/conv_one.py 3xtf32 128,32,32,32,3,1 > $O/ncu_wgrad.log 2>&1
fi
if has sgemm; then
# (5) SGEMM sweep (BASELINE config "sgemm") and one full capture of the TMA-fed GEMM kernel
python tools/sgemm_sweep.py 3xtf32 > $O/sgemm_3xtf32.txt 2>&1
python tools/sgemm_sweep.py tf32 > $O/sgemm_tf32.txt 2>&1
TNN_TMA_GEMM=0 python tools/sgemm_sweep.py 3xtf32 > $O/sgemm_3xtf32_regstaged.txt 2>&1
python tools/sgemm_sweep.py 3xtf32 4096 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_umma_gemm -c 1 -s 2 -o $O/r01_sgemm_nn4096 \
    python tools/sgemm_sweep.py 3xtf32 4096 > $O/ncu_sgemm.log 2>&1
python tools/conv_one.py 3xtf32 64,3,416,32,3,1 > $O/one_direct.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_conv_direct -c 1 -s 1 -o $O/r01_direct_c3f32 \
    python tools/conv_one.py 3xtf32 64,3,416,32,3,1 > $O/ncu_direct.log 2>&1
fi
ls -la $O
