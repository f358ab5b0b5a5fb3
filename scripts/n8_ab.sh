T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 8 --warmup 3 --no_stock --no_pipeline --no_e2e --no_cpu_baseline"
timeout 200 $T > gpurun_out/abc_a.json 2> gpurun_out/abc_a.err
DPC_DIRECT_NCCL=0 timeout 200 $T > gpurun_out/abc_b.json 2> gpurun_out/abc_b.err
DPC_WGRAD_DEFER=0 timeout 200 $T > gpurun_out/abc_c.json 2> gpurun_out/abc_c.err
DPC_DIRECT_NCCL=0 DPC_WGRAD_DEFER=0 timeout 200 $T > gpurun_out/abc_d.json 2> gpurun_out/abc_d.err
