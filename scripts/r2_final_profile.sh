#!/bin/bash
# Final 1-GPU evidence of the round: bench line, ncu launch list of one step, ncu --set full of its tensor-core launches.
#   gpurun --timeout 1500 -- 'bash scripts/r2_final_profile.sh'
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
timeout 200 python scripts/profile_one_step.py > gpurun_out/r02_one_step_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/r02_one_step_plain.log
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches.csv python scripts/profile_one_step.py > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --profile-from-start off --clock-control none -k regex:igemm --csv --metrics \
gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,launch__grid_size,launch__registers_per_thread,launch__shared_mem_per_block_dynamic \
    --log-file gpurun_out/r02_igemm_step_metrics.csv python scripts/profile_one_step.py > gpurun_out/r02_ncu_metrics.log 2>&1; echo "ncu metrics rc=$?"
# one --set full capture with source of the generator's 128->64 forward (the 11th fprop-type launch of the step)
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm_fprop -s 10 -c 1 -f \
    -o gpurun_out/r02_fprop_g128to64 python scripts/profile_one_step.py > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r02_*
