#!/bin/bash
# Round-end GPU recipe (run through gpurun from the repo root): GPU tests, the bench line, the ncu launch list of one step and
# ncu --set full captures of the dominant kernels.  Outputs land in gpurun_out/f_*; summaries are copied to profiles/ by hand.
mkdir -p gpurun_out; rm -f gpurun_out/f_status.log
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo pytest rc=$? >> gpurun_out/f_status.log
python bench.py > gpurun_out/f_bench_1gpu.json 2> gpurun_out/f_bench_1gpu.err; echo bench rc=$? >> gpurun_out/f_status.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/f_launches.csv python tests/prof_step.py 32 1 > gpurun_out/f_ncu_step.log 2>&1; echo ncu-list rc=$? >> gpurun_out/f_status.log
for w in fc1 fc2ln dgelu; do
  ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_bf16 --launch-skip 2 --launch-count 1 -f -o gpurun_out/f_gemm_$w python tests/gpu_gemm_one.py $w > gpurun_out/f_ncu_$w.log 2>&1; echo ncu-$w rc=$? >> gpurun_out/f_status.log
done
ncu --set full --clock-control none --import-source on --kernel-name regex:attn_bwd_tc_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/f_attn_bwd python tests/gpu_attn_bench.py > gpurun_out/f_ncu_attn_bwd.log 2>&1; echo ncu-attn rc=$? >> gpurun_out/f_status.log
cat gpurun_out/f_status.log; tail -3 gpurun_out/f_pytest.log
