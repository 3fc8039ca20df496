set -x
P=tools/probe/im2col_probe
$P 26 26 64 3 3 1 1 1 64 > gpurun_out/probe1.txt 2>&1
$P 26 26 64 3 3 1 2 1 64 >> gpurun_out/probe1.txt 2>&1
$P 13 13 96 20 3 1 1 1 128 >> gpurun_out/probe1.txt 2>&1
$P 52 52 32 2 1 0 1 1 32 >> gpurun_out/probe1.txt 2>&1
$P 8 8 16 2 3 1 1 1 16 >> gpurun_out/probe1.txt 2>&1
CLEAR_BIT21=1 $P 8 8 16 2 3 1 1 1 16 >> gpurun_out/probe1.txt 2>&1
$P 416 416 32 2 3 1 2 1 32 >> gpurun_out/probe1.txt 2>&1
cat gpurun_out/probe1.txt | grep -v "^   load" | tail -40
python tools/measure_peaks.py gpurun_out/r02_peaks.json
