set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_prefill -c 1 -o gpurun_out/attn_d64_split -f tools/bin/attn_test 128 512 12 12 64 0 1 > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
