# Round-2 evidence run (one B200): GPU suite, smoke, bench (ours + reference arm), then the ncu captures --
# each only after the same command exited 0 without ncu.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_ref.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum
python tools/kernel_probe.py > gpurun_out/r2_probe_plain.log 2>&1 && ncu --metrics $M --clock-control none -k regex:"legal|observation|apply|validate|replay|clone|reset|step" --csv --log-file gpurun_out/r2_api_kernels.csv python tools/kernel_probe.py > gpurun_out/r2_probe_ncu.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-kernels --no-configs > gpurun_out/r2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:playout_kernel -s 3 -c 1 -o gpurun_out/r2_playout_final python bench.py --steps 2 --warmup 3 --no-cpu --no-kernels --no-configs > gpurun_out/r2_ncu_playout.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-kernels --no-configs > gpurun_out/r2_plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-kernels --no-configs > gpurun_out/r2_ncu_launches.log 2>&1
python tools/kernel_probe.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:observation_kernel -s 2 -c 1 -o gpurun_out/r2_observation python tools/kernel_probe.py > gpurun_out/r2_ncu_obs.log 2>&1
tail -3 gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_smoke.log | tail -2; tail -c 400 gpurun_out/r2_bench.err
