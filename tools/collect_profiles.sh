#!/bin/bash
# Collect the evidence bench.py's numbers are explained by (run under gpurun, 1 GPU):
#   1. bench.py line (value / e2e / roofline / cpu_baseline)
#   2. ncu launch list of one mid-sequence decode step (per-launch device time, cold-cache, serialised)
#   3. ncu --set full of the dominant kernels (attention, gate|up contraction)
set -x
R=${1:-r01}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
tail -c 400 gpurun_out/${R}_bench.err
PG_STEPS=300 PG_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"^(gemm_tc|gemm_simt|attn_|resid_rmsnorm|swiglu|bias_act|cfg_sample|qkv_rope|embed_gather|decode_step)" -s 43300 -c 180 \
  --csv --log-file gpurun_out/${R}_launches_decode_step.csv python tools/profile_step.py > gpurun_out/p_a.log 2>&1
PG_STEPS=300 PG_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"attn_decode_v5" -s 7000 -c 1 \
  -o gpurun_out/${R}_attn_decode python tools/profile_step.py > gpurun_out/p_b.log 2>&1
PG_STEPS=300 PG_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 28900 -c 4 \
  -o gpurun_out/${R}_gemm_tc python tools/profile_step.py > gpurun_out/p_c.log 2>&1
PG_STEPS=2 PG_GRAPH=0 PG_VQ=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:"^(gemm_tc|gemm_simt|im2col|v_transpose|gn_|conv_epilogue|softmax_rows|vq_codebook|attn_prefill|qkv_rope|resid_rmsnorm|swiglu)" -c 2000 \
  --csv --log-file gpurun_out/${R}_launches_prefill_vq.csv python tools/profile_step.py > gpurun_out/p_d.log 2>&1
# 4. ncu --set full of the four prefill contractions of a layer (token tile 256): tensor-pipe utilisation
PG_STEPS=2 PG_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 1 -c 4 \
  -o gpurun_out/${R}_gemm_prefill python tools/profile_step.py > gpurun_out/p_e.log 2>&1
ls -la gpurun_out | tail -12
# afterwards, where ncu is installed:
#   python tools/summarize_launches.py gpurun_out/${R}_launches_decode_step.csv > profiles/${R}_launches_decode_step.txt
#   python tools/extract_ncu.py gpurun_out/${R}_attn_decode.ncu-rep > profiles/${R}_attn_decode.full.txt   (same for gemm_tc, gemm_prefill)
# 2-GPU line: gpurun --gpus 2 -- python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#   --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3
