#!/bin/bash
# gpurun --timeout 1500 -- "bash scripts/gpu_r2_ncu.sh": ncu --set full of the embedding kernels at cfg4 shapes (local and sharded-on-one-rank)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'emb_lookup|emb_scatter_kernel|emb_update' -s 48 -c 6 -f -o gpurun_out/prof_r02_large \
  python scripts/large_batch_steps.py cfg4 4 > gpurun_out/ncu_large.log 2>&1; echo "ncu large local rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'emb_update|p2p_grad_send|emb_lookup|p2p_unpack|p2p_route' -s 80 -c 5 -f -o gpurun_out/prof_r02_large_fs \
  python scripts/large_batch_steps.py cfg4 4 sharded > gpurun_out/ncu_large_fs.log 2>&1; echo "ncu large sharded rc=$?"
tail -3 gpurun_out/ncu_large.log gpurun_out/ncu_large_fs.log
ls -la gpurun_out/*.ncu-rep
