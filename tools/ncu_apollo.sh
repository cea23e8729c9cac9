#!/bin/bash
# Per-kernel counters of one Apollo restorer forward + MDX stft / istft (run under gpurun):  tools/ncu_apollo.sh <out prefix> [seconds]
out=${1:-gpurun_out/apollo}; S=${2:-10}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
M=$M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
M=$M,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed
M=$M,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
M=$M,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,lts__t_bytes.sum,lts__t_sector_hit_rate.pct
M=$M,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum
M=$M,sm__cycles_elapsed.avg,sm__cycles_elapsed.avg.per_second
python tools/run_apollo_once.py $S > ${out}_plain.log 2>&1 &&
ncu --metrics $M --clock-control none --profile-from-start off -o $out python tools/run_apollo_once.py $S > ${out}_ncu.log 2>&1 &&
ncu -i ${out}.ncu-rep --page raw --csv > ${out}_raw.csv
tail -n 2 ${out}_plain.log ${out}_ncu.log
