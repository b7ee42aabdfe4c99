# round 2: ncu --set full + per-role traces of the tensor-bound (>=64-channel) conv kernels (none existed in round 1)
set -x
O=gpurun_out/r2
mkdir -p $O
python profiles/layer_bench.py 32 > $O/layers_start.txt 2>&1 || exit 1
for spec in "0 256 256 3 14" "0 128 128 3 28" "0 64 64 3 56" "1 256 256 3 14" "1 256 128 3 28"; do
  tag=$(echo $spec | tr ' ' '_')
  python profiles/trace_one.py 32 $spec > $O/trace_$tag.txt 2>&1
done
LB="python profiles/layer_bench.py 32"
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 2 -c 1 -f -o $O/fprop_256_14 $LB 0 256 256 3 14 5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 2 -c 1 -f -o $O/fprop_128_28 $LB 0 128 128 3 28 5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 2 -c 1 -f -o $O/fprop_64_56 $LB 0 64 64 3 56 5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -s 2 -c 1 -f -o $O/wgrad_256_14 $LB 2 256 256 3 14 5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -s 2 -c 1 -f -o $O/wgrad_128_28 $LB 2 128 128 3 28 5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -s 2 -c 1 -f -o $O/wgrad_16_224 $LB 2 16 16 3 224 5 > /dev/null 2>&1
ls -la $O
