#!/bin/bash
# One multi-GPU gpurun call: bench.py under torchrun at N = 8 / 4 / 2 (sub-results c3, c4, one-call multi-GPU driver included),
# the concurrent PCIe probe at N = 8 / 4 / 2 / 1, and the GPU tests that need more than one device.
# usage (on the GPU box, from the repo root): bash tools/multi_gpu_run.sh OUTDIR "8 4 2"
out=${1:-gpurun_out/multi}; ns=${2:-"8 4 2"}
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
port=29500
for n in $ns; do
  port=$((port+1))
  devs=$(seq -s, 0 $((n-1)))
  CUDA_VISIBLE_DEVICES=$devs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 10 --warmup 3 > $out/bench_n$n.json 2> $out/bench_n$n.err
  echo "bench N=$n rc=$?"
  port=$((port+1))
  CUDA_VISIBLE_DEVICES=$devs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      tools/pcie_probe_multi.py > $out/pcie_n$n.txt 2> $out/pcie_n$n.err
  echo "pcie N=$n rc=$?"; cat $out/pcie_n$n.txt
done
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "two_contexts or multi_gpu or several_devices" 2>&1 | tail -4 | tee $out/pytest_multi.log
