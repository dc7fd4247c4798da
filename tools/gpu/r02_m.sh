#!/bin/bash
# r02 GPU call M: final validation — full GPU suite, final bench lines of every workload, ncu traffic of the z+ rule kernels
# and of the fused rollout, launch list
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -s > gpurun_out/m_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/m_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/m_bench_default.json 2> gpurun_out/m_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/m_bench_reference.json 2> gpurun_out/m_bench_reference.err
timeout 900 python bench.py --workload bert_base --steps 3 --warmup 3 --cpu-samples 3 > gpurun_out/m_bench_bert.json 2> gpurun_out/m_bench_bert.err
timeout 900 python bench.py --workload vit_large --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_vitl.json 2> gpurun_out/m_bench_vitl.err
timeout 900 python bench.py --workload deit_base_distilled --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_deit.json 2> gpurun_out/m_bench_deit.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-roofline --flags 3379 > gpurun_out/m_bench_bf16s1.json 2> gpurun_out/m_bench_bf16s1.err
timeout 600 python bench.py --scaling strong --batch 32 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/m_bench_b32.json 2> gpurun_out/m_bench_b32.err
timeout 900 ncu --set full --clock-control none -k regex:"te_tc_pair|abs_tf32" -c 4 -o gpurun_out/prof_pair_final python tools/profile_pair.py fc2 > gpurun_out/m_ncu_pair.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:rollout_row -c 1 -o gpurun_out/prof_rollout_final python tools/profile_kernels.py rollout_fused > gpurun_out/m_ncu_rollout.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/m_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/m_launches_bench.log 2>&1
