#!/bin/bash
# Round-2: full GPU suite, the default bench line (with extras), micro, ncu evidence (launch list + --set full).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -x > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/d_pytest.log | tail -8
timeout 900 python bench.py > gpurun_out/d_bench_default.json 2> gpurun_out/d_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --precision bf16 --no-extras --no-cpu-baseline > gpurun_out/d_bench_bf16.json 2> gpurun_out/d_bench_bf16.err; echo "bench bf16 rc=$?"
timeout 600 python bench.py --workload micro > gpurun_out/d_micro.json 2> gpurun_out/d_micro.err; echo "micro rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/d_ref.json 2> gpurun_out/d_ref.err; echo "ref rc=$?"
# launch list of two eager steps
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_waymo_fp32.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/d_ncu1.log 2>&1; echo "ncu list rc=$?"
# --set full: geometry kernels of one step (skip the two warm-up steps: 2 x 71 launches of ours incl. convs -> filter by name)
timeout 900 ncu --set full --clock-control none -k regex:"subm_probe|conv_insert|conv_rank|conv_nbr|group_|vox_|table_insert|fill_ranges|conv_small" -s 98 -c 49 -o gpurun_out/r2_geo_waymo_fp32 python profiles/run_geo.py --workload waymo_b4 > gpurun_out/d_ncu2.log 2>&1; echo "ncu geo rc=$?"
# --set full: conv layers 2 / 7 / 12 / 17 in both precisions (second of three repeats, after 2 x 20 warm-up launches)
for prec in fp32 bf16; do
  for layer in 2 7 12 17; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 41 -c 1 -o gpurun_out/r2_tc_${prec}_l$layer python profiles/run_layer.py --workload waymo_b4 --precision $prec --layer $layer > gpurun_out/d_ncu_${prec}_l$layer.log 2>&1
  done
done
echo done
