for cfg in "l2_ctas=1 l2_chains=14" "l2_ctas=2 l2_chains=10" "l2_ctas=2 l2_chains=14" "l2_ctas=3 l2_chains=14"; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__inst_executed.avg.per_cycle_elapsed,smsp__inst_executed.sum --clock-control none -k regex:k_compress_chain -s 1 -c 1 --csv python tools/prof_run.py 16384 0 smem_chains=0 spec_l2=16 $cfg 2>&1 | grep -E "k_compress_chain" | awk -F'","' '{print $(NF-2), $(NF)}' | tr '\n' ';'; echo " <= $cfg"
done
