#!/bin/bash
# ncu captures of the round-2 kernels (run on the GPU box: gpurun -- 'bash tools/capture_r02.sh TAG').
# Each kernel is captured once (--set full, one launch past the warm-up) after the same command ran clean without ncu.
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -c 1 --kernel-name-base demangled"
set -x
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-alt --no-cfg5 --no-pf > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || exit 1
# headline kernel, fp64 cfg2
$NCU -k regex:ffbs_fast_kernel --launch-skip 3 -o $OUT/${TAG}_fast_f64 -f \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-alt --no-cfg5 --no-pf > $OUT/${TAG}_ncu1.log 2>&1
# K = 10 block: fp64 then fp32 launches of ffbs_fastk_kernel
$NCU -k regex:ffbs_fastk_kernel --launch-skip 2 -o $OUT/${TAG}_fastk_f64 -f python tools/k10_probe.py > $OUT/${TAG}_ncu2.log 2>&1
$NCU -k "regex:ffbs_fastk_kernel<float" --launch-skip 2 -o $OUT/${TAG}_fastk_f32 -f python tools/k10_probe.py > $OUT/${TAG}_ncu3.log 2>&1
# fp32 cfg2 block (K <= 3 in fp32 runs on ffbs_warp_kernel)
$NCU -k "regex:ffbs_warp_kernel<float" --launch-skip 2 -o $OUT/${TAG}_warp_f32 -f \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-cfg5 --no-pf > $OUT/${TAG}_ncu4.log 2>&1
# launch list of the main loop
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_mainloop.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-alt --no-cfg5 --no-pf > $OUT/${TAG}_ncu5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_k10.csv \
  python tools/k10_probe.py > $OUT/${TAG}_ncu6.log 2>&1
ls -la $OUT/${TAG}_*
