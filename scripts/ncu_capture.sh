# ncu evidence for profiles/: launch list of two B=128 train steps + one `--set full` capture per kernel class.
#   gpurun --timeout 1500 -- 'bash scripts/ncu_capture.sh r2'      (then: python scripts/make_ncu_summary.py r2)
TAG=${1:-r2}
mkdir -p gpurun_out
P="python scripts/profile_step.py 128"
S="python scripts/profile_score.py 6144"
R="python scripts/profile_step.py 44 resnet34 224"
N="ncu --set full --clock-control none --import-source on -f"
# gpurun copies back at most 64 MiB: each report is reduced on the box to its raw-metric CSV (one row per metric) and the
# details page, then deleted; KEEP lists the reports worth keeping whole (source-level view)
KEEP="${KEEP:-}"
reduce() {
  for f in gpurun_out/${TAG}_*.ncu-rep; do
    [ -f "$f" ] || continue
    b=${f%.ncu-rep}
    ncu -i $f --page raw --csv > $b.raw.csv 2>/dev/null
    ncu -i $f --page details > $b.details.txt 2>/dev/null
    case " $KEEP " in *" $(basename $b) "*) ;; *) rm -f $f ;; esac
  done
}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv $P > gpurun_out/${TAG}_l.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_score_launches.csv $S > gpurun_out/${TAG}_sl.log 2>&1
# -s N: launches of that kernel to skip (0-based index of the captured launch inside the two steps)
timeout 200 $N -k regex:conv_tc_kernel        -s 3  -c 1 -o gpurun_out/${TAG}_conv_l3    $P > /dev/null 2>&1   # layer3.1.conv1 fwd
timeout 200 $N -k regex:conv_tc_halo          -s 8  -c 1 -o gpurun_out/${TAG}_conv_l1    $P > /dev/null 2>&1   # layer1.0.conv1 fwd (step 2)
timeout 200 $N -k regex:conv_tc_persist       -s 1  -c 1 -o gpurun_out/${TAG}_conv_l2    $P > /dev/null 2>&1   # layer2.0.conv2 fwd
timeout 200 $N -k regex:wgrad_tc_kernel       -s 10 -c 1 -o gpurun_out/${TAG}_wgrad_l3   $P > /dev/null 2>&1   # layer3.1.conv2 (after 5 head + 5 layer4 wgrads)
timeout 200 $N -k regex:wgrad_halo            -s 4  -c 1 -o gpurun_out/${TAG}_wgrad_l1   $P > /dev/null 2>&1
timeout 200 $N -k regex:stem_pool_fwd         -s 1  -c 1 -o gpurun_out/${TAG}_stem_fwd   $P > /dev/null 2>&1
timeout 200 $N -k regex:stem_pool_bwd_kernel  -s 1  -c 1 -o gpurun_out/${TAG}_stem_bwd   $P > /dev/null 2>&1
timeout 200 $N -k regex:bn_bwd_apply          -s 15 -c 1 -o gpurun_out/${TAG}_bn_bwd     $P > /dev/null 2>&1   # layer1 site
timeout 200 $N -k regex:head_chain_fwd        -s 1  -c 1 -o gpurun_out/${TAG}_head_fwd   $P > /dev/null 2>&1
timeout 200 $N -k regex:head_chain_bwd        -s 1  -c 1 -o gpurun_out/${TAG}_head_bwd   $P > /dev/null 2>&1
# score matmul + NCE (second iteration of profile_score.py)
timeout 200 $N -k regex:score_gemm_kernel     -s 1  -c 1 -o gpurun_out/${TAG}_score_fwd  $S > /dev/null 2>&1
timeout 200 $N -k regex:conv_tc_kernel        -s 1  -c 1 -o gpurun_out/${TAG}_score_dpred $S > /dev/null 2>&1
timeout 200 $N -k regex:wgrad_tc_kernel       -s 1  -c 1 -o gpurun_out/${TAG}_score_dfinf $S > /dev/null 2>&1
timeout 200 $N -k regex:ce_fwd                -s 1  -c 1 -o gpurun_out/${TAG}_ce_fwd     $S > /dev/null 2>&1
timeout 200 $N -k regex:ce_bwd                -s 1  -c 1 -o gpurun_out/${TAG}_ce_bwd     $S > /dev/null 2>&1
# on-device input pipeline (config-2 batch)
timeout 200 $N -k regex:augment_kernel        -s 3  -c 1 -o gpurun_out/${TAG}_augment    python scripts/bench_augment.py > /dev/null 2>&1
# R34 @ 224^2, B = 44 (BASELINE configs 4 / 5): the 14x14 layer3 site (56 % of that network's FLOPs)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_r34_launches.csv $R > gpurun_out/${TAG}_r34_l.log 2>&1
timeout 300 $N -k regex:conv_tc_kernel        -s 4  -c 1 -o gpurun_out/${TAG}_r34_conv_l3 $R > /dev/null 2>&1
reduce
ls -la gpurun_out | grep ${TAG}_ | head -80
