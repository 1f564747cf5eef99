#!/bin/bash
# 8-GPU box: bare pinned-copy ceiling at 1/2/4/8 ranks, then the bench line at N=8 and N=4
cd "$GRAFT_REPO_ROOT" || exit 1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/lscpu.txt 2>&1
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python profiles/scripts/pcie_probe.py 2>&1 | grep -E "ranks|softsplat_host"
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n profiles/scripts/pcie_probe.py 2>&1 | grep -E "ranks|softsplat_host"; fi
done | tee gpurun_out/pcie_probe_n.txt
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  echo "bench N=$n rc=$?"
done
cat gpurun_out/lscpu.txt; head -12 gpurun_out/topo.txt
