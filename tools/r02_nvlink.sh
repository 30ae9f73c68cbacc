#!/bin/bash
# usage (under gpurun --gpus 2): bash tools/r02_nvlink.sh — NVLink / DRAM bytes of the kernels that exchange over peer memory
# (rank 0 under ncu, rank 1 plain; plain launches because ncu cannot see kernel nodes of a graph with a conditional node)
mkdir -p gpurun_out
cat > /tmp/rank_wrap.sh <<'W'
#!/bin/bash
if [ "$LOCAL_RANK" = "0" ]; then
  exec ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes_srcunit_tex_aperture_peer.sum \
      -k regex:'walk2_kernel|subtree_push_kernel|sort_export_kernel|sort_fetch_kernel|finish_foreign_kernel' -c 30 --csv \
      --log-file gpurun_out/nvlink_rank0_$TAGN.csv python "$@"
else
  exec python "$@"
fi
W
chmod +x /tmp/rank_wrap.sh
for n in 1000000 10000000; do
  TAGN=$n KDNB_NO_GRAPH=1 timeout 400 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
      /tmp/rank_wrap.sh bench.py --gpus 2 --steps 2 --warmup 3 --number $n --no-10m > gpurun_out/nvlink_$n.log 2>&1
  echo "n=$n rc=$?"; grep -c walk2_kernel gpurun_out/nvlink_rank0_$n.csv; tail -3 gpurun_out/nvlink_rank0_$n.csv | cut -c1-300
done
