set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu"
$B > gpurun_out/b.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:skip_pool_bwd_kernel -s 3 -c 1 -f -o gpurun_out/g_skip $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:up_bwd_kernel -c 1 -f -o gpurun_out/g_upbwd $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:pool_act_kernel -c 1 -f -o gpurun_out/g_pool $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:pad_to_nhwc16 -c 1 -f -o gpurun_out/g_pad $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:dropout_bits -c 1 -f -o gpurun_out/g_drop $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:loss_ -c 2 -f -o gpurun_out/g_loss $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:upcat_kernel -s 3 -c 1 -f -o gpurun_out/g_upcat $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:bn_bwd_kernel -c 2 -f -o gpurun_out/g_bnbwd $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:bn_finalize -c 1 -f -o gpurun_out/g_fin $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:tc_wgrad_reduce -c 1 -f -o gpurun_out/g_red $B > /dev/null 2>&1
ls gpurun_out/*.ncu-rep
