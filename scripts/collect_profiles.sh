#!/bin/bash
# Round-end evidence run on one B200 (see profiles/): tests, smoke, bench, then the ncu passes of the SAME bench command.
# Every ncu pass starts only after the plain command has exited 0.  Outputs land in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 python -m pytest tests -m gpu -x -q > $O/t_all.log 2>&1; echo "pytest rc=$?" | tee -a $O/t_all.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 300 $BENCH > $O/bench_plain.json 2> $O/bench_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_bench.csv $BENCH > $O/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
# dram traffic of every GEMM launch of one timed step (3 warm-up steps x 113 GEMM launches are skipped)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_bf16 -s 339 -c 113 \
  --clock-control none --csv --log-file $O/gemm_dram.csv $BENCH > $O/ncu_dram.log 2>&1
echo "ncu dram rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 -s 341 -c 1 -o $O/prof_bench_gemm -f $BENCH > $O/ncu_g.log 2>&1
echo "ncu gemm full rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention_tc -s 84 -c 1 -o $O/prof_bench_attn -f $BENCH > $O/ncu_a.log 2>&1
echo "ncu attn full rc=$?"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/smi.txt
# configs[1] (base-224, batch 256) bench line and the memory-bound kernels: achieved GB/s (CUDA events), then one ncu capture each
timeout 300 python bench.py --workload base-224 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_base224.json 2> $O/bench_base224.err; echo "bench base rc=$?"
timeout 300 python scripts/kbench.py mem > $O/kbench_mem.txt 2>&1; echo "kbench mem rc=$?"
timeout 300 python scripts/kbench.py gemm attn > $O/kbench_gemm_attn.txt 2>&1; echo "kbench gemm/attn rc=$?"
timeout 400 ncu --set full --clock-control none -k regex:"freq_cols|freq_rows|map_attention|patchify_u8|head_fwd" -c 6 -o $O/prof_mem -f python scripts/prof_mem.py > $O/ncu_mem.log 2>&1
echo "ncu mem rc=$?"
