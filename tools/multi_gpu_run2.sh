#!/bin/bash
# second multi-GPU call: the concurrent PCIe probe at N = 8 / 4 / 2 / 1 and the end-to-end leg of bench.py at N = 8 for several
# sub-strip sizes.  usage: bash tools/multi_gpu_run2.sh OUTDIR
out=${1:-gpurun_out/multi2}; mkdir -p $out
port=29600
for n in 8 4 2 1; do
  port=$((port+1)); devs=$(seq -s, 0 $((n-1)))
  CUDA_VISIBLE_DEVICES=$devs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      tools/pcie_probe_multi.py > $out/pcie_n$n.txt 2> $out/pcie_n$n.err
  echo "pcie N=$n rc=$?"; cat $out/pcie_n$n.txt
done
for mb in 16 1000; do
  port=$((port+1))
  SRCNN_B200_STRIP_MB=$mb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus 8 --steps 10 --warmup 3 --no-sub > $out/bench_n8_strip$mb.json 2> $out/bench_n8_strip$mb.err
  echo "bench N=8 STRIP_MB=$mb rc=$?"
  python -c "import json; j=json.load(open('$out/bench_n8_strip$mb.json')); print('e2e', j['e2e']['value'])"
done
