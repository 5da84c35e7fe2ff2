NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/nccl_probe.py > gpurun_out/r2_nccl.log 2>&1
grep -E "GB/s|via |NVLS|P2P|SHM|Channel 00" gpurun_out/r2_nccl.log | head -20
nvidia-smi topo -m | head -8; df -h /dev/shm | tail -1
