#!/bin/bash
# Round 2, call I (one GPU): the evidence of the round with the final library -- bench lines of all five workloads
# (+ the reference arm), ncu launch lists, one full ncu capture per hot kernel, lag-off A/B.
mkdir -p gpurun_out
rm -f gpurun_out/r02_bench_lines.jsonl
for w in auni3d dendrite2d auni2d gg3d_hbsm pfhub1a; do
  timeout -k 5 400 python bench.py --workload $w >> gpurun_out/r02_bench_lines.jsonl 2> gpurun_out/r02_bench_$w.err
done
timeout -k 5 400 python bench.py --impl reference >> gpurun_out/r02_bench_lines.jsonl 2>> gpurun_out/r02_bench_auni3d.err
timeout -k 5 200 python bench.py --workload auni3d --no-lag --no-e2e --no-cpu-baseline >> gpurun_out/r02_bench_variants.jsonl 2>> gpurun_out/r02_bench_auni3d.err
timeout -k 5 200 python bench.py --workload auni3d --fd-flag 1 --no-e2e --no-cpu-baseline >> gpurun_out/r02_bench_variants.jsonl 2>> gpurun_out/r02_bench_auni3d.err
timeout -k 5 200 python bench.py --workload auni3d --newton warm --no-e2e --no-cpu-baseline --no-extras >> gpurun_out/r02_bench_variants.jsonl 2>> gpurun_out/r02_bench_auni3d.err
python - <<PY
import json
for f in ('gpurun_out/r02_bench_lines.jsonl','gpurun_out/r02_bench_variants.jsonl'):
  for l in open(f):
    if l.startswith('{'):
        d=json.loads(l)
        if d.get('impl')=='reference': print('REFERENCE', d['value'], d['cpu_baseline']['sample']); continue
        print(d['config']['workload'][:70], '| ms %.4f GCUPS %.2f frac %.3f e2e %s'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e'] and round(d['e2e']['value'],3)))
PY
for w in auni3d dendrite2d; do
  timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_$w.csv \
    python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_$w.log 2>&1
done
for w in auni3d dendrite2d auni2d gg3d_hbsm; do
  timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'rhs_|kks_' -s 8 -c 2 -f -o gpurun_out/prof_r02_$w \
    python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_full_r02_$w.log 2>&1
done
ls -la gpurun_out | grep r02_ | tail -20
