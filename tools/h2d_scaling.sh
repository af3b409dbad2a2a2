# Evidence for the end-to-end arm's multi-GPU ceiling: topology + concurrent pinned-copy bandwidth at N = 1, 2, 4, 8
OUT=gpurun_out/r2_h2d_scaling.txt
{
echo "## nvidia-smi topo -m"; nvidia-smi topo -m 2>&1
echo; echo "## lscpu (sockets / NUMA)"; lscpu 2>&1 | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread"
echo; echo "## numactl -H"; (numactl -H 2>&1 || echo "numactl not installed"); 
echo; echo "## /sys/devices/system/node"; for n in /sys/devices/system/node/node*; do echo "$n cpus $(cat $n/cpulist) mem $(grep MemTotal $n/meminfo | awk '{print $4, $5}')"; done
echo; echo "## PCIe link of every GPU"; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv 2>&1
echo; echo "## concurrent pinned copies, 1 GiB x 20 per rank (tools/h2d_scaling.py)"
for N in 1 2 4 8; do
  for B in 1 0; do
    BIND=$B python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N tools/h2d_scaling.py 2>/dev/null | tail -n1
  done
done
} > $OUT 2>&1
cat $OUT
