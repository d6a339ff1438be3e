M=dram__bytes_write.sum,dram__bytes_read.sum,lts__t_bytes.sum,gpu__time_duration.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_write_bytes.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed
for lean in 1 0; do
RNNT_B200_SAVE_HIDDEN=$lean ncu --metrics $M --clock-control none -k regex:joint_gemm_kernel -c 3 python scripts/time_fwd.py > gpurun_out/x_ncu_save$lean.txt 2>&1
done
