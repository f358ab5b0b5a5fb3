mkdir -p gpurun_out
P="python scripts/profile_step.py 128"
N="ncu --set full --clock-control none --import-source on -f"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r15_launches.csv $P > gpurun_out/r15_l.log 2>&1
timeout 200 $N -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/r15_conv_l3 $P > /dev/null 2>&1
timeout 200 $N -k regex:conv_tc_persist -s 0 -c 1 -o gpurun_out/r15_conv_l1 $P > /dev/null 2>&1
timeout 200 $N -k regex:wgrad_tc_kernel -s 6 -c 1 -o gpurun_out/r15_wgrad_l3 $P > /dev/null 2>&1
timeout 200 $N -k regex:stem_tc_fwd -s 0 -c 1 -o gpurun_out/r15_stem_fwd $P > /dev/null 2>&1
timeout 200 $N -k regex:stem_tail_bwd_apply -s 0 -c 1 -o gpurun_out/r15_stem_tail $P > /dev/null 2>&1
timeout 200 $N -k regex:bn_bwd_apply -s 15 -c 1 -o gpurun_out/r15_bn_bwd $P > /dev/null 2>&1
ls -la gpurun_out | grep r15
