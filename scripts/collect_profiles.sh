#!/bin/bash
# Round-end evidence run on one B200 (see profiles/): tests, smoke, bench, then the ncu passes of the SAME bench command.
# Every ncu pass starts only after the plain command has exited 0.  Outputs land in gpurun_out/ (prefix $1, default "r02").
set -u
O=gpurun_out
T=${1:-r02}
mkdir -p $O
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graphs 0"   # eager launches: ncu sees the same kernels, one by one
timeout 600 python bench.py --steps 5 --warmup 3 > $O/${T}_bench_so400m_b512.json 2> $O/${T}_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference_arm.json 2> $O/${T}_bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precise-residual 1 > $O/${T}_bench_so400m_b512_precise.json 2> $O/${T}_bench_precise.err; echo "precise rc=$?"
timeout 300 $BENCH > $O/${T}_bench_plain_for_ncu.json 2> $O/${T}_bench_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${T}_launches_bench.csv $BENCH > $O/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
# dram traffic of every GEMM launch of one timed step (3 warm-up steps x 113 GEMM launches are skipped)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_bf16 -s 339 -c 113 \
  --clock-control none --csv --log-file $O/${T}_gemm_dram.csv $BENCH > $O/ncu_dram.log 2>&1
echo "ncu dram rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 -s 341 -c 1 -o $O/${T}_prof_bench_gemm -f $BENCH > $O/ncu_g.log 2>&1
echo "ncu gemm full rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention_dq -s 84 -c 1 -o $O/${T}_prof_bench_attn -f $BENCH > $O/ncu_a.log 2>&1
echo "ncu attn full rc=$?"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/${T}_smi.txt
# configs[1] (base-224, batch 256) bench line, the auxiliary workloads and the memory-bound kernels
timeout 300 python bench.py --workload base-224 --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_base224_b256.json 2> $O/${T}_bench_base224.err; echo "bench base rc=$?"
for w in latency cifake head-train; do timeout 300 python bench.py --workload $w --steps 5 > $O/${T}_bench_$w.json 2> $O/${T}_bench_$w.err; echo "$w rc=$?"; done
timeout 300 python scripts/kbench.py mem > $O/${T}_kbench_mem.txt 2>&1; echo "kbench mem rc=$?"
timeout 300 python scripts/kbench.py gemm attn > $O/${T}_kbench_gemm_attn.txt 2>&1; echo "kbench gemm/attn rc=$?"
./scripts/ubench_softmax.bin > $O/${T}_ubench_softmax.txt 2>&1
