# One 8-GPU probe of the box: CPU / NUMA / PCIe topology, the concurrent host<->device copy roof, and the
# engine's own multi-GPU legs as they stand (torchrun weak scaling + the in-process sharding test).
O=gpurun_out/r02a; mkdir -p $O
(lscpu; echo; numactl -H 2>&1; echo; nvidia-smi topo -m; echo; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv; grep -m1 flags /proc/cpuinfo) > $O/box.txt 2>&1
tools/pcie_roof 8 > $O/pcie_roof.json 2> $O/pcie_roof.err
python -m pytest tests/test_gpu_parity.py -q -x -k in_process_multi_gpu > $O/pytest_inproc.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8.log 2>&1
tail -n 2 $O/pytest_inproc.log; tail -c 600 $O/bench_n8.log; cat $O/pcie_roof.json
