# ncu evidence for profiles/: launch list of two B=128 train steps + one `--set full` capture per kernel class.
#   gpurun --timeout 900 -- 'bash scripts/ncu_capture.sh r2'      (then: python scripts/make_ncu_summary.py r2)
TAG=${1:-r2}
mkdir -p gpurun_out
P="python scripts/profile_step.py 128"
N="ncu --set full --clock-control none --import-source on -f"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv $P > gpurun_out/${TAG}_l.log 2>&1
# -s N: launches of that kernel to skip (0-based index of the captured launch inside the two steps)
timeout 200 $N -k regex:conv_tc_kernel        -s 3  -c 1 -o gpurun_out/${TAG}_conv_l3    $P > /dev/null 2>&1   # layer3.1.conv1 fwd
timeout 200 $N -k regex:conv_tc_halo          -s 8  -c 1 -o gpurun_out/${TAG}_conv_l1    $P > /dev/null 2>&1   # layer1.0.conv1 fwd (step 2)
timeout 200 $N -k regex:conv_tc_persist       -s 1  -c 1 -o gpurun_out/${TAG}_conv_l2    $P > /dev/null 2>&1   # layer2.0.conv2 fwd
timeout 200 $N -k regex:wgrad_tc_kernel       -s 6  -c 1 -o gpurun_out/${TAG}_wgrad_l3   $P > /dev/null 2>&1   # layer3.1.conv2 wgrad
timeout 200 $N -k regex:wgrad_halo            -s 4  -c 1 -o gpurun_out/${TAG}_wgrad_l1   $P > /dev/null 2>&1   # layer1.1.conv2 wgrad (step 2)
timeout 200 $N -k regex:stem_s2d_fwd          -s 1  -c 1 -o gpurun_out/${TAG}_stem_fwd   $P > /dev/null 2>&1
timeout 200 $N -k regex:stem_s2d_wgrad        -s 1  -c 1 -o gpurun_out/${TAG}_stem_wgrad $P > /dev/null 2>&1
timeout 200 $N -k regex:stem_tail_bwd_apply   -s 0  -c 1 -o gpurun_out/${TAG}_stem_tail  $P > /dev/null 2>&1
timeout 200 $N -k regex:bn_bwd_apply          -s 15 -c 1 -o gpurun_out/${TAG}_bn_bwd     $P > /dev/null 2>&1   # layer1 site
ls -la gpurun_out | grep ${TAG}_
