#!/usr/bin/env bash
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0 SKIP_POPC=1 SKIP_GRM=1
timeout 500 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__data_pipe_lsu_wavefronts.sum,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_barrier.ratio --clock-control none -c 12 --csv --log-file gpurun_out/ts_launches.csv python tools/quick_king_bench.py 16384 65536 1 > gpurun_out/ts_ncu.log 2>&1
tail -3 gpurun_out/ts_ncu.log
