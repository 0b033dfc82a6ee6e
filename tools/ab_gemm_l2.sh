# DRAM traffic and time of the prefill gate/up GEMM under tile-order variants (ncu numbers are for traffic only)
for mb in ${MBS:-8 16 24 32 48}; do
  echo "== l2mb=$mb"
  HB_GEMM_L2MB=$mb tools/bin/gemm_test time 16384 28672 4096 ${EPI:-5} 256 | tail -1
  HB_GEMM_L2MB=$mb timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:gemm_tn -s 3 -c 1 tools/bin/gemm_test time 16384 28672 4096 ${EPI:-5} 256 2>&1 | grep -E "dram__|gpu__time|lts__"
done
