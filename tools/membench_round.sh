mkdir -p gpurun_out/membench
B=tools/bin/rcx_membench
$B > gpurun_out/membench/default.jsonl 2> gpurun_out/membench/default.err
$B quick -gran 32 > gpurun_out/membench/gran32.jsonl 2>&1
$B quick -gran 128 > gpurun_out/membench/gran128.jsonl 2>&1
ncu --metrics dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/membench/ncu_quick_default.csv $B quick > /dev/null 2>&1
ncu --metrics dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/membench/ncu_quick_gran32.csv $B quick -gran 32 > /dev/null 2>&1
cat gpurun_out/membench/default.jsonl
